"""A few launches of the voxel path at one shape (ncu target): python scripts/profile_voxel.py H W B N [norm]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cistaflow_b200 as cf  # noqa: E402
from cistaflow_b200 import synth  # noqa: E402

H, W, B, N = (int(a) for a in sys.argv[1:5])
norm = sys.argv[5] if len(sys.argv) > 5 else "std"
dev = torch.device("cuda", 0)
ev, off = synth.event_windows(B, N, H, W, 3)
ev, off = torch.from_numpy(ev).to(dev), torch.from_numpy(off).to(dev)
out = torch.empty((B, 5, H, W), device=dev)
for _ in range(3):
    cf.events_to_voxel_grid_batched(ev, off, 5, W, H, normalize=None if norm == "none" else norm,
                                    filter_hot_pixel=norm != "none", flavour="numpy", mode="atomic", out=out)
torch.cuda.synchronize()
print("ok")
