"""A few warm launches of the pyramid lookup at one shape (ncu target): python scripts/profile_lookup.py H W B"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cistaflow_b200 as cf  # noqa: E402
from cistaflow_b200 import synth  # noqa: E402

H, W, B = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (180, 240, 8)))
dev = torch.device("cuda", 0)
f1, f2, c0 = (torch.from_numpy(a).to(dev) for a in synth.corr_inputs(B, H, W, 3))
pyr = cf.build_pyramid(f1, f2, 4)
out = torch.empty((B, 324, c0.shape[2], c0.shape[3]), device=dev)
for _ in range(6):
    cf.corr_lookup(pyr, c0, 4, out=out)
torch.cuda.synchronize()
print("ok")
