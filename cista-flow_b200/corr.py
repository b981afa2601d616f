"""All-pairs correlation pyramid + lookup: B200 mirror of ``ERAFT/corr.py`` and
``DCEIFlow/core/corr/raft_corr.py`` (the two reference ``CorrBlock`` classes are
bit-identical).

``CorrBlock(fmap1, fmap2, num_levels=4, radius=4)`` builds the pyramid with the
tcgen05 TF32 GEMM (``cf_corr_build``), exposes it as ``corr_pyramid`` (list of
``[B*h*w, 1, h>>l, w>>l]`` float32 tensors, as in the reference) and
``__call__(coords)`` returns ``[B, num_levels*(2r+1)^2, h, w]`` through one
fused lookup launch (``cf_corr_lookup``) instead of the reference's ~40 small
kernels + 4 host->device copies per call.
"""
from __future__ import annotations

import os

import torch

from . import _lib

# 'tf32' (tcgen05 tensor cores, default), 'fp32' (SIMT, the reference's arithmetic)
DEFAULT_PRECISION = os.environ.get("CISTAFLOW_CORR_PRECISION", "tf32")
_PREC = {"tf32": _lib.CORR_TF32, "fp32": _lib.CORR_FP32, "3xtf32": _lib.CORR_3XTF32}


def coords_grid(batch, ht, wd, device=None):
    """ERAFT/utils.py:24-27 -- [B,2,ht,wd], channel 0 = x index, channel 1 = y index."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


def _prep(t: torch.Tensor, name: str) -> torch.Tensor:
    _lib.require_cuda(t, name)
    if torch.is_grad_enabled() and t.requires_grad:
        raise RuntimeError(f"cistaflow_b200 CorrBlock is inference-only: {name} requires grad")
    return t.float().contiguous()


def build_pyramid(fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4, precision: str | None = None,
                  out: list[torch.Tensor] | None = None):
    fmap1, fmap2 = _prep(fmap1, "fmap1"), _prep(fmap2, "fmap2")
    assert fmap1.shape == fmap2.shape and fmap1.dim() == 4
    B, D, h, w = fmap1.shape
    prec = precision or DEFAULT_PRECISION
    lib = _lib.load()
    if prec == "tf32" and (D % 32 != 0 or (h * w) % 4 != 0):
        prec = "fp32"  # shapes the tensor-core tiling does not cover (never the model's: D=256, h,w % 4 == 0)
    dev = fmap1.device
    shapes = [(B * h * w, 1, h >> l, w >> l) for l in range(num_levels)]
    if out is None:
        pyramid = [torch.empty(s, dtype=torch.float32, device=dev) for s in shapes]
    else:
        pyramid = list(out)
        assert len(pyramid) == num_levels
        for t, s in zip(pyramid, shapes):
            assert tuple(t.shape) == s and t.dtype == torch.float32 and t.is_contiguous() and t.device == dev
    with torch.cuda.device(dev):
        ws_bytes = lib.cf_corr_workspace_bytes(B, D, h, w, num_levels, _PREC[prec])
        ws = _lib.workspace(ws_bytes, dev)
        rc = lib.cf_corr_build(fmap1.data_ptr(), fmap2.data_ptr(), B, D, h, w, num_levels, _lib.pointer_array(pyramid),
                               _PREC[prec], _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev))
    _lib.check(rc, "cf_corr_build")
    return pyramid


def lookup(pyramid, coords: torch.Tensor, radius: int, out: torch.Tensor | None = None) -> torch.Tensor:
    coords = _prep(coords, "coords")
    B, two, h, w = coords.shape
    assert two == 2
    levels = len(pyramid)
    assert pyramid[0].shape[0] == B * h * w, "coords do not match the pyramid"
    k = 2 * radius + 1
    if out is None:
        out = torch.empty((B, levels * k * k, h, w), dtype=torch.float32, device=coords.device)
    lib = _lib.load()
    with torch.cuda.device(coords.device):
        rc = lib.cf_corr_lookup(_lib.pointer_array(pyramid), coords.data_ptr(), B, h, w, levels, radius, out.data_ptr(),
                                _lib.stream_ptr(coords.device))
    _lib.check(rc, "cf_corr_lookup")
    return out


class CorrBlock:
    """Drop-in for ``ERAFT/corr.py:12-60`` / ``raft_corr.py:15-65``."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, precision=None):
        self.num_levels = num_levels
        self.radius = radius
        with torch.no_grad():
            self.corr_pyramid = build_pyramid(fmap1, fmap2, num_levels, precision)

    def __call__(self, coords):
        with torch.no_grad():
            return lookup(self.corr_pyramid, coords, self.radius)

    @staticmethod
    def corr(fmap1, fmap2, precision=None):
        """[B, h, w, 1, h, w] = <fmap1, fmap2> / sqrt(D)  (ERAFT/corr.py:52-60)."""
        B, D, h, w = fmap1.shape
        with torch.no_grad():
            vol = build_pyramid(fmap1, fmap2, 1, precision)[0]
        return vol.view(B, h, w, 1, h, w)
