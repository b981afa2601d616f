"""Library-call restatement of the reference hot path (checker + CPU baseline).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Each function names the
reference lines it follows (paths relative to the reference checkout).  The
maths goes through the same NumPy / torch CPU calls the reference uses, so on
identical inputs these functions reproduce the reference bit for bit (checked
in ``tests/test_oracle_golden.py`` against fixtures generated from the
reference).

Differences from the reference that are deliberate:
  * inputs are never mutated (the reference rewrites ``events[:,3]`` /
    ``events[:,0]`` in place, utils/event_process.py:51,159);
  * ``preprocess_numpy`` returns float32 (under NumPy >= 2 the reference's
    result silently becomes float64 -- SURVEY.md F9 -- and every caller then
    feeds it to float32 convolutions).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# part 1: event stream -> voxel grid            (utils/event_process.py)
# --------------------------------------------------------------------------
def _normalised_time_np(t: np.ndarray, num_bins: int) -> np.ndarray:
    """utils/event_process.py:39-46 -- t* = (nb-1)(t-t0)/dT, dT==0 -> 1."""
    span = t[-1] - t[0]
    if span == 0:
        span = 1.0
    return (num_bins - 1) * (t - t[0]) / span


def voxel_grid_numpy(events: np.ndarray, num_bins: int, width: int, height: int) -> np.ndarray:
    """``events_to_voxel_grid`` (utils/event_process.py:15-72), is_reverse=False.

    fp64 weights, ``np.add.at`` into a float32 grid: every addition is carried
    out in fp64 and rounded to fp32, strictly in event order, all "left"
    contributions before all "right" ones.
    """
    assert events.shape[1] == 4 and num_bins > 0 and width > 0 and height > 0
    grid = np.zeros(num_bins * height * width, np.float32)
    if len(events) == 0:
        return grid.reshape(num_bins, height, width)
    ev = np.array(events, dtype=np.float64, copy=True)
    tn = _normalised_time_np(ev[:, 0], num_bins)
    col = ev[:, 1].astype(np.uint)
    row = ev[:, 2].astype(np.uint)
    sgn = ev[:, 3].copy()
    sgn[sgn == 0] = -1
    lo = tn.astype(np.uint)
    frac = tn - lo
    w_lo = sgn * (1.0 - frac)
    w_hi = sgn * frac
    plane = width * height
    keep = lo < num_bins
    np.add.at(grid, col[keep] + row[keep] * width + lo[keep] * plane, w_lo[keep])
    keep = (lo + 1) < num_bins
    np.add.at(grid, col[keep] + row[keep] * width + (lo[keep] + 1) * plane, w_hi[keep])
    return grid.reshape(num_bins, height, width)


def voxel_grid_pol_numpy(events: np.ndarray, num_bins: int, width: int, height: int) -> np.ndarray:
    """``events_to_voxel_grid_pol`` (utils/event_process.py:75-123): one plane
    per (bin, polarity), all weights positive."""
    assert events.shape[1] == 4 and num_bins > 0 and width > 0 and height > 0
    grid = np.zeros(num_bins * 2 * height * width, np.float32)
    if len(events) == 0:
        return grid.reshape(num_bins, 2, height, width)
    ev = np.array(events, dtype=np.float64, copy=True)
    tn = _normalised_time_np(ev[:, 0], num_bins)
    col = ev[:, 1].astype(np.uint)
    row = ev[:, 2].astype(np.uint)
    chan = ev[:, 3].astype(np.uint)
    mag = ev[:, 3].copy()
    mag[mag == 0] = 1.0
    lo = tn.astype(np.uint)
    frac = tn - lo
    w_lo = mag * (1.0 - frac)
    w_hi = mag * frac
    plane = width * height
    keep = lo < num_bins
    np.add.at(grid, col[keep] + row[keep] * width + chan[keep] * plane + lo[keep] * plane * 2, w_lo[keep])
    keep = (lo + 1) < num_bins
    np.add.at(grid, col[keep] + row[keep] * width + chan[keep] * plane + (lo[keep] + 1) * plane * 2, w_hi[keep])
    return grid.reshape(num_bins, 2, height, width)


def voxel_grid_torch(events: torch.Tensor, num_bins: int, width: int, height: int) -> torch.Tensor:
    """``events_to_voxel_grid_pytorch`` (utils/event_process.py:127-190).

    On a CPU fp64 event tensor this is the BIT-EXACT oracle for the
    deterministic binning mode: fp64 time normalisation, fp32 weights,
    ``index_add_`` = sequential fp32 accumulation in event order (left pass,
    then right pass).
    """
    assert events.shape[1] == 4 and num_bins > 0 and width > 0 and height > 0
    with torch.no_grad():
        grid = torch.zeros(num_bins * height * width, dtype=torch.float32, device=events.device)
        if len(events) == 0:
            return grid.view(num_bins, height, width)
        ev = events.clone()
        span = ev[-1, 0] - ev[0, 0]
        if span == 0:
            span = 1.0
        tn = (num_bins - 1) * (ev[:, 0] - ev[0, 0]) / span
        col = ev[:, 1].long()
        row = ev[:, 2].long()
        sgn = ev[:, 3].float()
        sgn[sgn == 0] = -1
        lo = torch.floor(tn)
        lo_i = lo.long()
        frac = tn - lo
        w_lo = sgn * (1.0 - frac.float())
        w_hi = sgn * frac.float()
        plane = width * height
        keep = (lo < num_bins) & (lo >= 0)
        grid.index_add_(0, col[keep] + row[keep] * width + lo_i[keep] * plane, w_lo[keep])
        keep = ((lo + 1) < num_bins) & (lo >= 0)
        grid.index_add_(0, col[keep] + row[keep] * width + (lo_i[keep] + 1) * plane, w_hi[keep])
    return grid.view(num_bins, height, width)


def mvsec_voxel_torch(xs, ys, ts, ps, num_bins: int, height: int, width: int) -> torch.Tensor:
    """``events_to_voxel_torch(..., temporal_bilinear=True)`` (data_readers/MVSEC_utils.py:253-303): one
    ``index_put_(accumulate=True)`` per bin with weights ``ps * max(0, 1 - |t* - bin|)``, t* = (ts - ts[0]) / dT * (B-1).
    The reference's ``torch.max(zeros_f32, f64)`` promotes to fp64, the product with the (integer or float) polarity
    stays fp64 and is cast to fp32 for the accumulation (``events_to_image_torch``, :243-250)."""
    xs, ys, ts, ps = (torch.as_tensor(a) for a in (xs, ys, ts, ps))
    with torch.no_grad():
        grid = torch.zeros((num_bins, height, width), dtype=torch.float32)
        span = ts[-1] - ts[0]
        tn = (ts - ts[0]) / span * (num_bins - 1)
        rows, cols = ys.long(), xs.long()
        for b in range(num_bins):
            tri = torch.clamp(1.0 - torch.abs(tn - b), min=0.0)
            grid[b].index_put_((rows, cols), (ps * tri).float(), accumulate=True)
    return grid


def mvsec_binary_search(t: torch.Tensor, l: int, r: int, x) -> int:
    """``binary_search_torch_tensor`` (MVSEC_utils.py:184-201), side='left'."""
    while l <= r:
        mid = l + (r - l) // 2
        midval = t[mid]
        if midval == x:
            return mid
        if midval < x:
            l = mid + 1
        else:
            r = mid - 1
    return l


def mvsec_voxel_naive_torch(xs, ys, ts, ps, num_bins: int, height: int, width: int) -> torch.Tensor:
    """``events_to_voxel_torch(..., temporal_bilinear=False)`` (MVSEC_utils.py:292-300): bin bi accumulates the polarities of
    the events with ts in [ts[0] + dt*bi, ts[0] + dt*(bi+1)), dt = ts[-1] - ts[0] (the whole window, as written there)."""
    xs, ys, ts, ps = (torch.as_tensor(a) for a in (xs, ys, ts, ps))
    with torch.no_grad():
        grid = torch.zeros((num_bins, height, width), dtype=torch.float32)
        dt = ts[-1] - ts[0]
        for b in range(num_bins):
            tstart = ts[0] + dt * b
            tend = tstart + dt
            beg = mvsec_binary_search(ts, 0, len(ts) - 1, tstart)
            end = mvsec_binary_search(ts, 0, len(ts) - 1, tend)
            grid[b].index_put_((ys[beg:end].long(), xs[beg:end].long()), ps[beg:end].float(), accumulate=True)
    return grid


def mvsec_events_to_voxel(events_xytp: np.ndarray, num_bins: int, height: int, width: int,
                          event_polarity: bool = False) -> np.ndarray:
    """``eventsToVoxel`` (MVSEC_utils.py:384-403) on rows (x, y, t, p): ``eventsToXYTP(process=True)`` (:348-364)
    narrows x, y, p to int32 and normalises t to [0, 1] in NumPy fp64 first; with ``event_polarity`` the positive
    (p > 0) and non-positive events are binned with unit weights into separate grids, concatenated (:306-343)."""
    xs = events_xytp[:, 0].astype(np.int32)
    ys = events_xytp[:, 1].astype(np.int32)
    ps = events_xytp[:, 3].astype(np.int32)
    ts = events_xytp[:, 2]
    ts = (ts - ts[0]) / (ts[-1] - ts[0])
    if not event_polarity:
        return mvsec_voxel_torch(xs, ys, ts, ps, num_bins, height, width).numpy()
    pt = torch.from_numpy(ps)
    pos = mvsec_voxel_torch(xs, ys, ts, torch.where(pt > 0, 1.0, 0.0), num_bins, height, width)
    neg = mvsec_voxel_torch(xs, ys, ts, torch.where(pt <= 0, 1.0, 0.0), num_bins, height, width)
    return torch.cat([pos, neg], 0).numpy()


def preprocess_numpy(grid: np.ndarray, mode: str = "std", filter_hot_pixel: bool = False) -> np.ndarray:
    """``event_preprocess`` (utils/event_process.py:193-216); hot-pixel
    threshold 25/num_bins.  Returns float32 (see module docstring)."""
    g = np.array(grid, copy=True)
    nb = g.shape[0]
    if filter_hot_pixel:
        g[abs(g) > 25.0 / nb] = 0
    if mode == "maxmin":
        g = (g - g.min()) / (g.max() - g.min() + 1e-8)
    elif mode == "std":
        nz = g != 0
        cnt = nz.sum()
        if cnt > 0:
            mean = g.sum() / cnt
            std = np.sqrt((g ** 2).sum() / cnt - mean ** 2)
            g = nz.astype(np.float32) * (g - mean) / (std + 1e-8)
    else:
        raise AssertionError("mode must be 'maxmin' or 'std'")
    return np.asarray(g, dtype=np.float32)


def preprocess_torch(grid: torch.Tensor, mode: str = "std", filter_hot_pixel: bool = False) -> torch.Tensor:
    """``event_preprocess_pytorch`` (utils/event_process.py:219-239); hot-pixel
    threshold 20/num_bins (sic -- differs from the NumPy variant)."""
    g = grid.clone()
    nb = g.shape[0]
    if filter_hot_pixel:
        g[abs(g) > 20.0 / nb] = 0
    if mode == "maxmin":
        g = (g - g.min()) / (g.max() - g.min() + 1e-8)
    elif mode == "std":
        nz = g != 0
        cnt = nz.sum()
        if cnt > 0:
            mean = g.sum() / cnt
            std = torch.sqrt((g ** 2).sum() / cnt - mean ** 2)
            g = nz.float() * (g - mean) / (std + 1e-8)
    return g


# --------------------------------------------------------------------------
# part 2: flow-guided warp                        (utils/flow_utils.py)
# --------------------------------------------------------------------------
def warp(img: torch.Tensor, flow: torch.Tensor, mode: str = "forward") -> torch.Tensor:
    """``forwardWarp.forward`` / ``backWarp.forward``
    (utils/flow_utils.py:153-190 / 83-120).  Both are a bilinear *gather*
    (``grid_sample``, align_corners=True, reflection padding) at
    (x -/+ u, y -/+ v) with the reference's ``2*(x/W - 0.5)`` normalisation.
    """
    hgt, wid = img.shape[-2:]
    gx, gy = np.meshgrid(np.arange(wid), np.arange(hgt))
    gx = torch.tensor(gx, device=flow.device)
    gy = torch.tensor(gy, device=flow.device)
    u, v = flow[:, 0], flow[:, 1]
    if mode == "forward":
        x = gx.unsqueeze(0).expand_as(u).float() - u
        y = gy.unsqueeze(0).expand_as(v).float() - v
    else:
        x = gx.unsqueeze(0).expand_as(u).float() + u
        y = gy.unsqueeze(0).expand_as(v).float() + v
    x = 2 * (x / wid - 0.5)
    y = 2 * (y / hgt - 0.5)
    return F.grid_sample(img, torch.stack((x, y), dim=3), align_corners=True, padding_mode="reflection")


def downsample_flow(flow: torch.Tensor) -> torch.Tensor:
    """e2v/e2v_model.py:190 -- x0.5 bilinear, align_corners=True, flow VALUES
    are not rescaled."""
    return F.interpolate(flow, scale_factor=0.5, mode="bilinear", align_corners=True)


def warp_frame_and_codes(img, codes, flow, mode="forward"):
    """The per-frame warp step of e2v/e2v_model.py:188-191 (and :240-243)."""
    return warp(img, flow, mode), warp(codes, downsample_flow(flow), mode)


def upflow8_unpad(flow_lr: torch.Tensor, pad_h: int, pad_w: int) -> torch.Tensor:
    """DCEIFlow/utils/sample_utils.py:66-68 (``8 * F.interpolate(size=8x, bilinear, align_corners=True)``) followed by
    ``ImagePadder.unpad`` (utils/image_process.py:103-107: the padding sits on the top/left): DCEIFlow/DCEIFlow.py:222-227."""
    h, w = flow_lr.shape[-2:]
    up = 8 * F.interpolate(flow_lr, size=(8 * h, 8 * w), mode="bilinear", align_corners=True)
    return up[..., pad_h:, pad_w:]


def warp_frame_and_codes_upflow8(img, codes, flow_lr, pad_h: int, pad_w: int, mode="forward"):
    """upflow8 + unpad + the per-frame warp step (e2v/e2v_model.py:188-191): (warped image, warped codes, flow_final)."""
    flow = upflow8_unpad(flow_lr, pad_h, pad_w).contiguous()
    wi, wz = warp_frame_and_codes(img, codes, flow, mode)
    return wi, wz, flow


# --------------------------------------------------------------------------
# part 3: all-pairs correlation + pyramid lookup   (ERAFT/corr.py == DCEIFlow/core/corr/raft_corr.py)
# --------------------------------------------------------------------------
def coords_grid(batch: int, ht: int, wd: int) -> torch.Tensor:
    """ERAFT/utils.py:24-27 -- channel 0 = x index, channel 1 = y index."""
    ys, xs = torch.meshgrid(torch.arange(ht), torch.arange(wd), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


def corr_volume(fmap1: torch.Tensor, fmap2: torch.Tensor) -> torch.Tensor:
    """``CorrBlock.corr`` (ERAFT/corr.py:52-60): <f1[:,i], f2[:,j]> / sqrt(D),
    shape [B, h, w, 1, h, w]."""
    b, d, h, w = fmap1.shape
    vol = torch.matmul(fmap1.view(b, d, h * w).transpose(1, 2), fmap2.view(b, d, h * w))
    return vol.view(b, h, w, 1, h, w) / torch.sqrt(torch.tensor(d).float())


def corr_pyramid(fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4) -> list[torch.Tensor]:
    """``CorrBlock.__init__`` (ERAFT/corr.py:13-27): level l is
    [B*h*w, 1, h>>l, w>>l] (avg_pool2d floors odd sizes)."""
    vol = corr_volume(fmap1, fmap2)
    b, h, w, d, h2, w2 = vol.shape
    vol = vol.reshape(b * h * w, d, h2, w2)
    pyr = [vol]
    for _ in range(num_levels - 1):
        vol = F.avg_pool2d(vol, 2, stride=2)
        pyr.append(vol)
    return pyr


def _pixel_sampler(img: torch.Tensor, xy: torch.Tensor) -> torch.Tensor:
    """``bilinear_sampler`` (ERAFT/utils.py:7-21): grid_sample in pixel
    coordinates, align_corners=True, zero padding."""
    hgt, wid = img.shape[-2:]
    xg, yg = xy.split([1, 1], dim=-1)
    xg = 2 * xg / (wid - 1) - 1
    yg = 2 * yg / (hgt - 1) - 1
    return F.grid_sample(img, torch.cat([xg, yg], dim=-1), align_corners=True)


def corr_lookup(pyramid: list[torch.Tensor], coords: torch.Tensor, radius: int = 4) -> torch.Tensor:
    """``CorrBlock.__call__`` (ERAFT/corr.py:29-50).  Note the window is
    transposed: ``delta = stack(meshgrid(dy, dx))`` is added to (x, y), so
    output channel l*(2r+1)^2 + i*(2r+1) + j samples at (x + i - r, y + j - r)."""
    r = radius
    xy = coords.permute(0, 2, 3, 1)
    b, h, w, _ = xy.shape
    span = torch.linspace(-r, r, 2 * r + 1)
    delta = torch.stack(torch.meshgrid(span, span, indexing="ij"), dim=-1).to(coords.device)
    delta = delta.view(1, 2 * r + 1, 2 * r + 1, 2)
    out = []
    for lvl, vol in enumerate(pyramid):
        centre = xy.reshape(b * h * w, 1, 1, 2) / 2 ** lvl
        out.append(_pixel_sampler(vol, centre + delta).view(b, h, w, -1))
    return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def voxel_flow_warp(voxel: torch.Tensor, displacement: torch.Tensor, reverse_time: bool = False):
    """loss.py:27-83 (voxel_warping_flow_loss) through the same library calls: one grid_sample per channel
    over the whole grid, channel i kept.  Returns (variance, summed [N,1,H,W], warped [N,C,H,W])."""
    if reverse_time:
        displacement = -displacement
    n, c, h, w = voxel.shape
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    xx, yy = xx.float(), yy.float()
    dx, dy = displacement[:, 0], displacement[:, 1]
    inc = 1.0 / (c - 1.0)
    summed = torch.zeros((n, 1, h, w), dtype=voxel.dtype)
    warped = torch.zeros((n, c, h, w), dtype=voxel.dtype)
    for i in range(c):
        ratio = (1.0 - i * inc) if reverse_time else i * inc
        grid = torch.stack([xx + dx * ratio, yy + dy * ratio], dim=3)
        grid[:, :, :, 1] = (2.0 * grid[:, :, :, 1]) / h - 1.0
        grid[:, :, :, 0] = (2.0 * grid[:, :, :, 0]) / w - 1.0
        ch = F.grid_sample(voxel, grid, align_corners=True)[:, i:i + 1]
        summed += ch
        warped[:, i:i + 1] = ch
    return summed.var(), summed, warped


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    """loss.py:15-24 -- 20*log10(1/sqrt(mse)), 100 when mse < 1e-10."""
    mse = float(torch.mean((a.double() - b.double()) ** 2))
    return 100.0 if mse < 1e-10 else 20.0 * math.log10(1.0 / math.sqrt(mse))
