"""Per-CTA timeline of corr_lookup_kernel (experiment build with -DCF_TRACE): python scripts/lookup_trace.py H W B"""
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
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.cf_trace_buffer_lookup.argtypes = [ctypes.c_void_p]
f1, f2, c0 = synth.corr_inputs(B, H, W, 3)
f1, f2, c0 = (torch.from_numpy(a).to(dev) for a in (f1, f2, c0))
pyr = cf.build_pyramid(f1, f2, 4)
out = torch.empty((B, 324, c0.shape[2], c0.shape[3]), device=dev)
for _ in range(3):
    cf.corr_lookup(pyr, c0, 4, out=out)
torch.cuda.synchronize()
slots, ncta = 256, 8192
buf = torch.zeros(ncta * slots, dtype=torch.int64, device=dev)
assert lib.cf_trace_buffer_lookup(buf.data_ptr()) == 0
if len(sys.argv) > 4 and sys.argv[4] == "warm":   # pyramid L2-resident, as inside a frame's 12 lookups
    for _ in range(3):
        cf.corr_lookup(pyr, c0, 4, out=out)
else:
    torch.empty(64 << 20, dtype=torch.float32, device=dev).fill_(1.0)  # flush L2
torch.cuda.synchronize()
cf.corr_lookup(pyr, c0, 4, out=out)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(ncta, slots)
sm = t[t[:, 3] > 0][:, 4] - 1
t = t[t[:, 3] > 0][:, :4]
t0 = t[:, 0].min()
print(f"{len(t)} CTAs; CTA start: median {np.median(t[:, 0] - t0) / 1e3:.2f} max {(t[:, 0].max() - t0) / 1e3:.2f} us; "
      f"end: median {np.median(t[:, 3] - t0) / 1e3:.2f} max {(t[:, 3].max() - t0) / 1e3:.2f} us")
seg = (t - t[:, :1]) / 1e3
print("per-CTA medians since CTA start (us): gathered %.2f, barrier passed %.2f, stored %.2f" % tuple(np.median(seg[:, 1:], axis=0)))
print("per-CTA 95th pct                     : gathered %.2f, barrier passed %.2f, stored %.2f" % tuple(np.percentile(seg[:, 1:], 95, axis=0)))

# per-SM view: does the tail follow the number of CTAs an SM received?
import collections
per_sm = collections.defaultdict(list)
for k in range(len(t)):
    per_sm[int(sm[k])].append((t[k, 3] - t0) / 1e3)
by_count = collections.defaultdict(list)
for s_, ends in per_sm.items():
    by_count[len(ends)].append(max(ends))
for n in sorted(by_count):
    v = np.array(by_count[n])
    print(f"SMs with {n} CTA(s): {len(v):3d}; last CTA end median {np.median(v):.2f} max {v.max():.2f} us")
slow = sorted(per_sm.items(), key=lambda kv: -max(kv[1]))[:8]
print("slowest SMs:", ", ".join(f"sm{s_}:{len(e)}cta:{max(e):.2f}" for s_, e in slow))
