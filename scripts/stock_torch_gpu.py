#!/usr/bin/env python
"""The comparator SURVEY.md section 8(d) asks for beside the CPU baseline: the SAME hot-path ops on the same B200 through
stock PyTorch CUDA kernels, i.e. what the unmodified reference launches when its tensors live on the GPU
(`index_add_` voxeliser per window, `F.grid_sample` warps, `torch.matmul` + `avg_pool2d` pyramid, `grid_sample`
lookups).  It is a measurement arm only: nothing in the package imports it, and it imports nothing from `oracle/`.

The op sequences restate (they do not import) the reference's:
  voxel      utils/event_process.py:125-188 (events_to_voxel_grid_pytorch) + :219-240 (event_preprocess_pytorch)
  warp       utils/flow_utils.py:122-190 (forwardWarp) / :40-120 (backWarp), e2v/e2v_model.py:188-191 (frame + codes)
  pyramid    ERAFT/corr.py:13-27, lookup ERAFT/corr.py:29-50, ERAFT/utils.py:8-21 (bilinear_sampler)

Standalone:  python scripts/stock_torch_gpu.py [--workload configs[4]] [--out profiles/r02/stock_torch_gpu.json]
bench.py calls `measure()` for its `stock_torch_gpu` key (skipped with --no-extra).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- the reference's op sequences, on whatever device the inputs live ---------------------------------------
def voxel_window(events, nb, W, H):
    """One window, events [N,4] fp64 (t,x,y,p) on the device; the caller's tensor is overwritten like the reference's."""
    grid = torch.zeros(nb * H * W, dtype=torch.float32, device=events.device)
    if len(events) == 0:
        return grid.view(nb, H, W)
    t0, t1 = events[0, 0], events[-1, 0]
    span = t1 - t0
    if span == 0:                                  # host read-back, as in the reference
        span = 1.0
    events[:, 0] = (nb - 1) * (events[:, 0] - t0) / span
    ts = events[:, 0]
    xs, ys = events[:, 1].long(), events[:, 2].long()
    pol = events[:, 3].float()
    pol[pol == 0] = -1
    tis = torch.floor(ts)
    til = tis.long()
    frac = (ts - tis).float()
    ok = (tis < nb) & (tis >= 0)
    grid.index_add_(0, xs[ok] + ys[ok] * W + til[ok] * W * H, (pol * (1.0 - frac))[ok])
    ok = ((tis + 1) < nb) & (tis >= 0)
    grid.index_add_(0, xs[ok] + ys[ok] * W + (til[ok] + 1) * W * H, (pol * frac)[ok])
    return grid.view(nb, H, W)


def preprocess_window(g):
    nb = g.shape[0]
    g[abs(g) > 20.0 / nb] = 0
    nz = g != 0
    n = nz.sum()
    if n > 0:                                      # host read-back, as in the reference
        mean = g.sum() / n
        std = torch.sqrt((g ** 2).sum() / n - mean ** 2)
        g = nz.float() * (g - mean) / (std + 1e-8)
    return g


def voxel_step(events, offsets_host, nb, W, H):
    out = []
    for b in range(len(offsets_host) - 1):
        out.append(preprocess_window(voxel_window(events[offsets_host[b]:offsets_host[b + 1]], nb, W, H)))
    return torch.stack(out)


class StockWarp:
    def __init__(self, W, H, sign):
        self.gx, self.gy = np.meshgrid(np.arange(W), np.arange(H))
        self.W, self.H, self.sign = W, H, sign

    def __call__(self, img, flow):
        gx = torch.tensor(self.gx, device=flow.device)          # per-call upload, as in the reference
        gy = torch.tensor(self.gy, device=flow.device)
        u, v = flow[:, 0], flow[:, 1]
        x = gx.unsqueeze(0).expand_as(u).float() + self.sign * u
        y = gy.unsqueeze(0).expand_as(v).float() + self.sign * v
        grid = torch.stack((2 * (x / self.W - 0.5), 2 * (y / self.H - 0.5)), dim=3)
        return F.grid_sample(img, grid, align_corners=True, padding_mode="reflection")


def warp_step(img, codes, flow, warps):
    wi = warps[0](img, flow)
    half = F.interpolate(flow, scale_factor=0.5, mode="bilinear", align_corners=True)   # values not rescaled
    wz = warps[1](codes, half)
    return wi, wz


def pyramid_step(f1, f2, levels):
    B, D, h, w = f1.shape
    corr = torch.matmul(f1.view(B, D, h * w).transpose(1, 2), f2.view(B, D, h * w)).view(B, h, w, 1, h, w)
    corr = corr / torch.sqrt(torch.tensor(D).float())
    corr = corr.reshape(B * h * w, 1, h, w)
    pyr = [corr]
    for _ in range(levels - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyr.append(corr)
    return pyr


def lookup_step(pyr, coords, r):
    coords = coords.permute(0, 2, 3, 1)
    B, h, w, _ = coords.shape
    outs = []
    for i, corr in enumerate(pyr):
        d = torch.linspace(-r, r, 2 * r + 1)
        delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).to(coords.device)
        c = coords.reshape(B * h * w, 1, 1, 2) / 2 ** i + delta.view(1, 2 * r + 1, 2 * r + 1, 2)
        Hl, Wl = corr.shape[-2:]
        xg, yg = c.split([1, 1], dim=-1)
        grid = torch.cat([2 * xg / (Wl - 1) - 1, 2 * yg / (Hl - 1) - 1], dim=-1)
        outs.append(F.grid_sample(corr, grid, align_corners=True).view(B, h, w, -1))
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float()


# ---- timing ---------------------------------------------------------------------------------------------------
def _time(fn, reps, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps        # ms; eager launches + the reference's own host read-backs included


def measure(d, cfg, ours=None, reps=3):
    """d: one device input set of bench.HotPath (events, offsets, img, codes, flow, fmap1, fmap2, coords[list]).
    ours: optional (vox, lookup0, wi, wz) from the library on the same set -> max|diff| per op is reported.
    Returns ms per op for the whole batch and the frames/s of the resulting step."""
    nb, W, H, r, L = cfg["bins"], cfg["W"], cfg["H"], cfg["radius"], cfg["levels"]
    B = d["img"].shape[0]
    off = d["offsets"].tolist()
    sign = -1.0 if cfg["warp_mode"] == "forward" else 1.0
    warps = (StockWarp(W, H, sign), StockWarp(W // 2, H // 2, sign))
    res, diff = {}, {}
    with torch.no_grad():
        copies = [d["events"].clone() for _ in range(reps + 1)]     # the voxeliser overwrites its input column 0
        it = iter(copies)
        res["voxel_bin+normalise"] = _time(lambda: voxel_step(next(it), off, nb, W, H), reps)
        del copies, it
        res["warp_frame_and_codes"] = _time(lambda: warp_step(d["img"], d["codes"], d["flow"], warps), reps)
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            res["corr_build" + ("[allow_tf32]" if tf32 else "")] = _time(lambda: pyramid_step(d["fmap1"], d["fmap2"], L), reps)
        torch.backends.cuda.matmul.allow_tf32 = False
        pyr = pyramid_step(d["fmap1"], d["fmap2"], L)
        notes = {}
        try:
            res["corr_lookup"] = _time(lambda: lookup_step(pyr, d["coords"][0], r), reps)
        except RuntimeError as e:
            # grid_sample dispatches to cuDNN's spatial-transformer sampler, which rejects the [B*N,1,h,w] batch of a
            # many-stream step; the reference as written stops here.  Time torch's native sampler instead and say so.
            notes["corr_lookup"] = ("stock dispatch (cuDNN grid sampler) failed on this batch: " + str(e)[:120]
                                    + " -- timed with torch.backends.cudnn.enabled=False")
            torch.backends.cudnn.enabled = False
            res["corr_lookup"] = _time(lambda: lookup_step(pyr, d["coords"][0], r), reps)
        if ours is not None:
            vox, look, wi, wz = ours
            sv = voxel_step(d["events"].clone(), off, nb, W, H)
            swi, swz = warp_step(d["img"], d["codes"], d["flow"], warps)
            sl = lookup_step(pyr, d["coords"][0], r)
            diff = {"voxel_bin+normalise": float((sv - vox).abs().max()), "warp_frame": float((swi - wi).abs().max()),
                    "warp_codes": float((swz - wz).abs().max()),
                    "corr_lookup_rel": float((sl - look).abs().max() / sl.abs().max())}
            # triangulate the warp on stream 0: the same torch ops on the CPU (what the parity tests pin the library to)
            cw = (StockWarp(W, H, sign), StockWarp(W // 2, H // 2, sign))
            ci, cz = warp_step(d["img"][:1].cpu(), d["codes"][:1].cpu(), d["flow"][:1].cpu(), cw)
            diff["warp_codes_stream0"] = {"stock_gpu_vs_stock_cpu": float((swz[:1].cpu() - cz).abs().max()),
                                          "library_vs_stock_cpu": float((wz[:1].cpu() - cz).abs().max())}
            diff["warp_frame_stream0"] = {"stock_gpu_vs_stock_cpu": float((swi[:1].cpu() - ci).abs().max()),
                                          "library_vs_stock_cpu": float((wi[:1].cpu() - ci).abs().max())}
        del pyr
    torch.backends.cudnn.enabled = True
    torch.cuda.empty_cache()
    step_ms = (res["voxel_bin+normalise"] + res["warp_frame_and_codes"] + res["corr_build"]
               + cfg["lookups"] * res["corr_lookup"])
    return {"what": "the reference's own torch op sequences for this path on the same GPU (stock PyTorch CUDA kernels, "
                    "eager, fp32 matmul as the reference runs it; its per-window host read-backs included)",
            "ms": res, "step_ms": step_ms, "frames_per_s": B / step_ms * 1e3,
            "notes": notes or None, "max_abs_diff_vs_library": diff or None, "streams": B, "torch": torch.__version__}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="configs[4]")
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    sys.path.insert(0, ROOT)
    import bench
    cfg = bench.make_cfg(args.workload)
    B = args.streams or cfg["streams"]
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    hp = bench.HotPath(cfg, B, dev, 1234, n_sets=1)
    d = hp.sets[0]
    cf = hp.cf
    vox = cf.events_to_voxel_grid_batched(d["events"], d["offsets"], cfg["bins"], cfg["W"], cfg["H"], normalize="std",
                                          filter_hot_pixel=True, flavour="torch", mode="atomic")   # 20/nb threshold, like the op timed
    blk = cf.CorrBlock(d["fmap1"], d["fmap2"], num_levels=cfg["levels"], radius=cfg["radius"], precision="fp32")
    look = blk(d["coords"][0])
    wi, wz = cf.warp_frame_and_codes(d["img"], d["codes"], d["flow"], cfg["warp_mode"])
    del blk
    out = measure(d, cfg, ours=(vox, look, wi, wz))
    out["workload"] = cfg["workload"]
    text = json.dumps(out, indent=1)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
