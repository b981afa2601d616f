"""Print the clock64 timeline of CTA 0 of the tcgen05 correlation kernel (debug stamps)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cistaflow_b200 as cf
from cistaflow_b200 import synth

dev = torch.device("cuda", 0)
lib = cf.load_library()
for (H, W, B) in ((180, 240, 8), (480, 640, 1)):
    f1, f2, _ = synth.corr_inputs(B, H, W, seed=1)
    a, b = torch.from_numpy(f1).to(dev), torch.from_numpy(f2).to(dev)
    for rep in range(3):
        cf.build_pyramid(a, b, 4)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 32)()
    lib.cf_debug_tc_timeline(buf)
    t0 = buf[0]
    names = {0: "start", 1: "setup done", 2: "first TMA issued", 19: "last commit issued", 20: "epilogue: accumulator ready",
             21: "epilogue done", 22: "kernel end"}
    names.update({3 + k: f"stage kb={k} landed" for k in range(8)})
    print(f"--- {H}x{W} B={B}")
    for k in sorted(names):
        if buf[k]:
            print(f"  {names[k]:32s} +{buf[k] - t0:8d} cycles")
