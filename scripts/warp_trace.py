"""Per-CTA timeline of warp_persist_kernel (experiment build with -DCF_TRACE, see build.build_variant).

    python cista-flow_b200/build.py   # normal build
    python -c "import sys; sys.path.insert(0,'cista-flow_b200'); import build; build.build_variant('trace', ['CF_TRACE'])"
    CISTAFLOW_LIB=build/libcistaflow_trace.so python scripts/warp_trace.py
"""
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

H, W, B = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (180, 240, 8)))
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.cf_trace_buffer.argtypes = [ctypes.c_void_p]
slots = lib.cf_trace_slots()
img, codes, flow = (torch.from_numpy(a).to(dev) for a in synth.warp_inputs(B, H, W, 3, 128, flow_kind="smooth"))
for _ in range(3):
    cf.warp_frame_and_codes(img, codes, flow, "forward")
torch.cuda.synchronize()
buf = torch.zeros(148 * slots, dtype=torch.int64, device=dev)
assert lib.cf_trace_buffer(buf.data_ptr()) == 0
# flush L2
junk = torch.empty(64 << 20, dtype=torch.float32, device=dev).fill_(1.0)
torch.cuda.synchronize()
cf.warp_frame_and_codes(img, codes, flow, "forward")
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(148, slots).astype(np.int64)
t0 = t[:, 0][t[:, 0] > 0].min()
print(f"kernel {_lib.load().cf_last_kernel().decode()}  {H}x{W} B={B}")
ends = t[:, 2] - t0
print(f"CTA start spread {(t[:, 0].max() - t0) / 1e3:.2f} us; image done (median) {np.median(t[:, 1] - t0) / 1e3:.2f} us; "
      f"CTA end min/median/max {ends.min() / 1e3:.2f}/{np.median(ends) / 1e3:.2f}/{ends.max() / 1e3:.2f} us")
for cta in (0, 73, 147):
    r = t[cta]
    print(f"--- CTA {cta}: start {(r[0] - t0) / 1e3:.2f} image_done {(r[1] - t0) / 1e3:.2f} end {(r[2] - t0) / 1e3:.2f}")
    n = 0
    while 8 + 4 * n + 3 < slots and r[8 + 4 * n] > 0:
        iss, w0, w1, done = (r[8 + 4 * n + k] - t0 for k in range(4))
        print(f"  chunk {n:3d}: issue {iss / 1e3:7.2f}  wait_begin {w0 / 1e3:7.2f}  full {w1 / 1e3:7.2f} (load latency "
              f"{(w1 - iss) / 1e3:5.2f}, stalled {(w1 - w0) / 1e3:5.2f})  done {done / 1e3:7.2f} (compute {(done - w1) / 1e3:5.2f})")
        n += 1
