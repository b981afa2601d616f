"""FWL metric: B200 mirror of the reference's ``loss.voxel_warping_flow_loss`` (loss.py:27-83).

Same name, arguments and return convention, so that ``from loss import voxel_warping_flow_loss``
(test_wo_flow.py:20,161; test_mvsec.py:180) can be re-pointed here by ``install()``.
"""
from __future__ import annotations

import torch

from . import _lib


def voxel_warping_flow_loss(voxel: torch.Tensor, displacement: torch.Tensor, output_images: bool = False,
                            reverse_time: bool = False):
    """Variance of the flow-warped, channel-summed voxel grid.

    voxel [N,C,H,W], displacement [N,2,H,W] (CUDA, float32) -> 0-dim float32 CUDA tensor
    (+ {'voxel_grid', 'voxel_grid_warped'} when ``output_images``), like the reference.
    One launch instead of C grid_sample passes over the whole grid; no host sync."""
    _lib.require_cuda(voxel, "voxel")
    _lib.require_cuda(displacement, "displacement")
    with torch.no_grad():   # the reference detaches every warped channel (loss.py:66-67)
        v = voxel.detach().float().contiguous()
        d = displacement.detach().float().contiguous()
        N, C, H, W = v.shape
        assert d.shape == (N, 2, H, W), "displacement must be [N,2,H,W]"
        dev = v.device
        summed = torch.empty((N, 1, H, W), dtype=torch.float32, device=dev)
        warped = torch.empty_like(v) if output_images else None
        mean_var = torch.empty(2, dtype=torch.float64, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            ws_bytes = lib.cf_voxel_flow_warp_workspace_bytes(N, H, W)
            ws = _lib.workspace(ws_bytes, dev)
            rc = lib.cf_voxel_flow_warp(v.data_ptr(), d.data_ptr(), N, C, H, W, int(bool(reverse_time)), _lib.ptr(warped),
                                        summed.data_ptr(), mean_var.data_ptr(), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev))
        _lib.check(rc, "cf_voxel_flow_warp")
        tc_loss = mean_var[1].to(v.dtype)
    if output_images:
        return tc_loss, {"voxel_grid": voxel, "voxel_grid_warped": warped}
    return tc_loss
