"""Seeded synthetic inputs for the three sub-paths.

Shared by the parity tests, ``smoke()`` and ``bench.py`` so that the CUDA path,
the oracle and the CPU baseline always see identical data.  NumPy only; callers
move the arrays to the device.

Shapes follow the reference's conventions:
  events  float64 [N, 4] rows (t, x, y, p), t sorted, p in {0, 1}
          (data_readers/event_readers.py:15-20)
  img     float32 [B, 1, H, W]   previous reconstruction
  codes   float32 [B, 128, H/2, W/2]  CISTA-LSTC sparse codes Z (e2v/e2v_model.py:28-32)
  flow    float32 [B, 2, H, W]   ch0 = u (x), ch1 = v (y)
  fmap    float32 [B, 256, Hp/8, Wp/8] with Hp, Wp = H, W rounded up to x32
          (utils/image_process.py:70-101)
"""
from __future__ import annotations

import numpy as np

T_BASE = 1000.0     # seconds: absolute stamps that do not fit fp32
T_SPAN = 0.03


def seed_for(config: int, frame: int = 0) -> int:
    return 1234 + 1000 * config + frame


def events(n: int, height: int, width: int, seed: int, hot_fraction: float = 0.2,
           hot_pixels: float = 0.01) -> np.ndarray:
    """One window of ``n`` events; ``hot_fraction`` of them land on
    ``hot_pixels`` of the sensor (stresses atomic contention)."""
    rng = np.random.default_rng(seed)
    t = T_BASE + np.sort(rng.random(n)) * T_SPAN
    x = rng.integers(0, width, n)
    y = rng.integers(0, height, n)
    n_hot_px = max(1, int(hot_pixels * height * width))
    hot = rng.integers(0, height * width, n_hot_px)
    pick = rng.random(n) < hot_fraction
    sel = hot[rng.integers(0, n_hot_px, n)]
    x = np.where(pick, sel % width, x)
    y = np.where(pick, sel // width, y)
    p = rng.integers(0, 2, n)
    return np.stack([t, x, y, p], axis=1).astype(np.float64)


def event_windows(batch: int, n: int, height: int, width: int, seed: int):
    """``batch`` windows concatenated + int64 offsets [batch+1] (the batched
    boundary form, include/cistaflow.h ``cf_voxel_bin``)."""
    wins = [events(n, height, width, seed + 7919 * b) for b in range(batch)]
    offsets = np.zeros(batch + 1, np.int64)
    offsets[1:] = np.cumsum([len(w) for w in wins])
    return np.concatenate(wins, axis=0), offsets


def soft_shrink(x: np.ndarray, lam: float) -> np.ndarray:
    return np.sign(x) * np.maximum(np.abs(x) - lam, 0.0)


def smooth_field(rng, batch: int, channels: int, height: int, width: int, sigma: float, cell: int = 16) -> np.ndarray:
    """Low-frequency random field: N(0, sigma^2) on a coarse grid (one node per
    ``cell`` pixels), bilinearly interpolated -- neighbouring pixels move together,
    like real optical flow."""
    ch, cw = max(2, height // cell + 1), max(2, width // cell + 1)
    coarse = sigma * rng.standard_normal((batch, channels, ch, cw))
    ys = np.linspace(0, ch - 1, height)
    xs = np.linspace(0, cw - 1, width)
    y0 = np.minimum(ys.astype(np.int64), ch - 2)
    x0 = np.minimum(xs.astype(np.int64), cw - 2)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    c00 = coarse[:, :, y0][:, :, :, x0]
    c01 = coarse[:, :, y0][:, :, :, x0 + 1]
    c10 = coarse[:, :, y0 + 1][:, :, :, x0]
    c11 = coarse[:, :, y0 + 1][:, :, :, x0 + 1]
    return ((1 - fy) * ((1 - fx) * c00 + fx * c01) + fy * ((1 - fx) * c10 + fx * c11)).astype(np.float32)


def warp_inputs(batch: int, height: int, width: int, seed: int, code_channels: int = 128,
                flow_sigma: float = 5.0, flow_kind: str = "noise"):
    """(img, codes, flow): img ~ U(0,1); codes ~ soft-shrunk N(0,1) (about 60 %
    zeros, like real sparse codes).  flow_kind:
      'noise'   independent N(0, sigma^2) px per pixel + an affine component: the
                adversarial case for the gather (every tap of a warp in a
                different cache line) -- used by the parity tests;
      'smooth'  what a flow network produces: the nets predict flow at 1/8
                resolution and up-sample it, so neighbouring pixels move together.
                Calibrated on the flow_final tensors recorded from the reference
                models (tests/golden/trace_*.npz: std 1.4-2.1 px, max ~9 px, mean
                |d flow/dx| 0.06 px/px): N(0, 2^2) px nodes every 32 px, bilinearly
                interpolated, + the affine component + 0.02 px jitter -- used by bench.py."""
    rng = np.random.default_rng(seed)
    img = rng.random((batch, 1, height, width), dtype=np.float32)
    codes = soft_shrink(rng.standard_normal((batch, code_channels, height // 2, width // 2),
                                            dtype=np.float32), 0.84).astype(np.float32)
    yy, xx = np.meshgrid(np.linspace(-1, 1, height, dtype=np.float32),
                         np.linspace(-1, 1, width, dtype=np.float32), indexing="ij")
    smooth = np.stack([3.0 * xx - 2.0 * yy, 1.5 * yy + 2.5 * xx])[None]
    if flow_kind == "smooth":
        flow = smooth_field(rng, batch, 2, height, width, 0.4 * flow_sigma, cell=32) + smooth \
            + 0.02 * rng.standard_normal((batch, 2, height, width), dtype=np.float32)
    else:
        flow = flow_sigma * rng.standard_normal((batch, 2, height, width), dtype=np.float32) + smooth
    return img, codes, flow.astype(np.float32)


def padded_dims(height: int, width: int, multiple: int = 32):
    return -(-height // multiple) * multiple, -(-width // multiple) * multiple


def corr_inputs(batch: int, height: int, width: int, seed: int, dim: int = 256,
                coord_sigma: float = 3.0):
    """(fmap1, fmap2, coords) at the 1/8-resolution of the x32-padded image."""
    hp, wp = padded_dims(height, width)
    h, w = hp // 8, wp // 8
    rng = np.random.default_rng(seed)
    f1 = rng.standard_normal((batch, dim, h, w), dtype=np.float32)
    f2 = rng.standard_normal((batch, dim, h, w), dtype=np.float32)
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    grid = np.stack([xs, ys])[None]
    coords = (grid + coord_sigma * rng.standard_normal((batch, 2, h, w), dtype=np.float32)).astype(np.float32)
    return f1, f2, coords
