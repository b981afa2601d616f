"""Per-CTA timelines of the tiled voxel path (experiment build with -DCF_TRACE).

    python -c "import sys; sys.path.insert(0,'cista-flow_b200'); import build; build.build_variant('trace', ['CF_TRACE'])"
    python scripts/voxel_trace.py H W B N      # prints pass A (partition) and pass B (tiles) timelines
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

H, W, B, N = (int(a) for a in sys.argv[1:5])
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.cf_trace_buffer_voxel.argtypes = [ctypes.c_void_p]
ev, off = synth.event_windows(min(B, 4), N, H, W, 3)
rep = -(-B // min(B, 4))
ev = np.concatenate([ev] * rep)[: B * N]
off = np.arange(B + 1, dtype=np.int64) * N
ev, off = torch.from_numpy(ev).to(dev), torch.from_numpy(off).to(dev)
out = torch.empty((B, 5, H, W), device=dev)


def run():
    cf.events_to_voxel_grid_batched(ev, off, 5, W, H, normalize="std", filter_hot_pixel=True, flavour="numpy",
                                    mode="atomic_tiled", out=out)


for _ in range(3):
    run()
torch.cuda.synchronize()
slots, ncta = 256, 16384
buf = torch.zeros(ncta * slots, dtype=torch.int64, device=dev)
assert lib.cf_trace_buffer_voxel(buf.data_ptr()) == 0
torch.empty(64 << 20, dtype=torch.float32, device=dev).fill_(1.0)  # flush L2
torch.cuda.synchronize()
run()
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(ncta, slots).astype(np.float64)
a = t[t[:, 204] > 0][:, 200:205]          # pass A: start | hist zeroed | binned | sorted | written (last chunk of the CTA)
t0 = a[:, 0].min()
print(f"{H}x{W} B={B} N={N}")
print(f"pass A: {len(a)} CTAs; first start 0.00, last start {(a[:, 0].max() - t0) / 1e3:.2f} us, last end {(a[:, 4].max() - t0) / 1e3:.2f} us")
seg = a[:, 1:5] - a[:, 1:2]
print("  last chunk of a CTA, us since its hist-zeroed stamp (median): binned %.2f, sorted %.2f, written %.2f; CTA lifetime median %.2f us"
      % (*(np.median(seg, axis=0)[1:] / 1e3), np.median(a[:, 4] - a[:, 0]) / 1e3))
b = t[(t[:, 0] > 0) & (t[:, 204] == 0)]   # pass B CTAs (1-D grid: their ids overlap pass A's x = 0 column only when B == 1)
if len(b) == 0:
    b = t[t[:, 0] > 0]
names = ["ACC start", "ACC: adds done + published", "FIN: took the buffer", "FIN: window complete", "FIN: written"]
print(f"pass B: {len(b)} CTAs; us relative to pass A's first start")
r = 0
while 8 * r + 4 < 200 and (b[:, 8 * r] > 0).any():
    live = b[b[:, 8 * r + 4] > 0]
    s = (live[:, 8 * r: 8 * r + 5] - t0) / 1e3
    d = np.diff(s, axis=1)
    print(f"  round {r:2d}: {len(live):3d} CTAs, start med {np.median(s[:, 0]):7.2f}; phase durations med/max: "
          + "; ".join(f"{n} {np.median(d[:, k]):.2f}/{d[:, k].max():.2f}" for k, n in enumerate(names[1:]))
          + f"; end max {s[:, 4].max():.2f}")
    r += 1
