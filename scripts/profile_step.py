"""Short eager run of the configs[1] hot-path step for ncu (no graphs, no e2e)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cistaflow_b200 as cf

cfg = dict(bench.CFG)
dev = torch.device("cuda", 0)
sets = bench.make_host_inputs(cfg, 3234)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
import numpy as np
dsets = []
for s in sets:
    d = {k: torch.from_numpy(v).to(dev) for k, v in s.items() if isinstance(v, np.ndarray)}
    d["coords"] = [torch.from_numpy(c).to(dev) for c in s["coords"]]
    dsets.append(d)
for it in range(iters):
    d = dsets[it % len(dsets)]
    vox = cf.events_to_voxel_grid_batched(d["events"], d["offsets"], cfg["bins"], cfg["W"], cfg["H"], normalize="std",
                                          filter_hot_pixel=True, flavour="numpy", mode="atomic")
    blk = cf.CorrBlock(d["fmap1"], d["fmap2"], num_levels=cfg["levels"], radius=cfg["radius"])
    outs = [blk(c) for c in d["coords"]]
    wi, wz = cf.warp_frame_and_codes(d["img"], d["codes"], d["flow"], cfg["warp_mode"])
torch.cuda.synchronize()
print("profile_step done", tuple(vox.shape), tuple(outs[-1].shape), tuple(wz.shape))  # (no torch kernels in the capture)
