"""A few launches of cf.warp_frame_and_codes at one shape (ncu target): python scripts/profile_warp.py H W B"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cistaflow_b200 as cf  # noqa: E402
from cistaflow_b200 import synth  # noqa: E402

H, W, B = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (180, 240, 8)))
dev = torch.device("cuda", 0)
img, codes, flow = (torch.from_numpy(a).to(dev) for a in synth.warp_inputs(B, H, W, 3, 128, flow_kind="smooth"))
for _ in range(4):
    cf.warp_frame_and_codes(img, codes, flow, "forward")
torch.cuda.synchronize()
print("ok")
