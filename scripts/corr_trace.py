"""Per-tile timeline of corr_tc_kernel (experiment build with -DCF_TRACE): python scripts/corr_trace.py H W B [cta ...]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CISTAFLOW_LIB", os.path.join(ROOT, "build", "libcistaflow_trace.so"))
import cistaflow_b200 as cf  # noqa: E402
from cistaflow_b200 import _lib, synth  # noqa: E402

H, W, B = (int(a) for a in sys.argv[1:4])
ctas = [int(a) for a in sys.argv[4:]] or [0, 1, 100]
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.cf_trace_buffer_corr.argtypes = [ctypes.c_void_p]
f1, f2, _ = synth.corr_inputs(B, H, W, 3)
f1, f2 = (torch.from_numpy(a).to(dev) for a in (f1, f2))
for _ in range(3):
    cf.build_pyramid(f1, f2, 4)
torch.cuda.synchronize()
slots, ncta = 256, 256
buf = torch.zeros(ncta * slots, dtype=torch.int64, device=dev)
assert lib.cf_trace_buffer_corr(buf.data_ptr()) == 0
cf.build_pyramid(f1, f2, 4)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(ncta, slots)
t0 = t[:, 16:][t[:, 16:] > 0].min()
print(f"CF_TC_FLAGS={os.environ.get('CF_TC_FLAGS', '0')}  {H}x{W} B={B}; us since the first stamp of the launch")
for c in ctas:
    print(f"--- CTA {c}")
    for k in range(15):
        r = t[c, 16 + 16 * k: 32 + 16 * k]
        if r[12] == 0 and r[2] == 0 and r[0] == 0:
            break
        u = lambda v: (v - t0) / 1e3 if v else float('nan')
        land = " ".join(f"{u(v):6.2f}" for v in r[3:5])
        w4 = ""
        if r[7]:
            done = {0: r[14], 1: r[5], 2: r[6], 3: r[15], 4: r[7], 5: r[8], 6: r[9], 7: r[10]}
            w4 = " || epilogue warps done " + " ".join(f"w{w}:{u(v):6.2f}" for w, v in done.items())
        print(f"  tile {k:2d}: loads {u(r[0]):6.2f}..{u(r[1]):6.2f} | acc free {u(r[2]):6.2f} landed [{land}] commit {u(r[11]):6.2f} "
              f"| epi ready {u(r[12]):6.2f} l0 issued {u(r[13]):6.2f} done {u(r[14]):6.2f}{w4}")
