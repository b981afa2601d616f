"""Per-CTA timelines of the tiled voxel path (experiment build with -DCF_TRACE).

    python -c "import sys; sys.path.insert(0,'cista-flow_b200'); import build; build.build_variant('trace', ['CF_TRACE'])"
    python scripts/voxel_trace.py H W B N [A|B]      # which pass to trace (they share the buffer)
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
which = sys.argv[5] if len(sys.argv) > 5 else "B"
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.cf_trace_buffer_voxel.argtypes = [ctypes.c_void_p]
ev, off = synth.event_windows(B, N, H, W, 3)
ev, off = torch.from_numpy(ev).to(dev), torch.from_numpy(off).to(dev)
out = torch.empty((B, 5, H, W), device=dev)


def run():
    cf.events_to_voxel_grid_batched(ev, off, 5, W, H, normalize="std", filter_hot_pixel=True, flavour="numpy",
                                    mode="atomic", out=out)


for _ in range(3):
    run()
torch.cuda.synchronize()
slots, ncta = 256, 4096
buf = torch.zeros(ncta * slots, dtype=torch.int64, device=dev)
assert lib.cf_trace_buffer_voxel(buf.data_ptr()) == 0
torch.empty(64 << 20, dtype=torch.float32, device=dev).fill_(1.0)  # flush L2
torch.cuda.synchronize()
run()
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(ncta, slots)
# both passes wrote into the same slots; pass B ran last and overwrote slots 0..5 of its CTAs (grid <= 148)
names_b = ["start", "bounds staged", "runs added", "partial published", "stats combined", "written"]
names_a = ["start", "window located", "events binned", "sorted in smem", "written"]
if which == "B":
    nb_cta = int((t[:, 5] > 0).sum())
    tb = t[:148]
    live = tb[:, 0] > 0
    t0 = tb[live, 0].min()
    waves = int((tb[live][0, ::8][: slots // 8] > 0).sum())
    print(f"pass B: {live.sum()} CTAs, {waves} wave(s); times in us relative to the first CTA start")
    for wv in range(waves):
        seg = tb[live][:, 8 * wv: 8 * wv + 6] - t0
        med = np.median(seg, axis=0) / 1e3
        mx = seg.max(axis=0) / 1e3
        print(f"  wave {wv}: " + "; ".join(f"{n} {m:.2f} (max {x:.2f})" for n, m, x in zip(names_b, med, mx)))
else:
    ta = t[t[:, 204] > 0][:, 200:205]   # pass A stamps live in slots 200..204
    t0 = ta[:, 0].min()
    seg = ta[:, :5] - ta[:, :1]
    print(f"pass A: {len(ta)} CTAs; start spread {(ta[:, 0].max() - t0) / 1e3:.2f} us; end max {(ta[:, 4].max() - t0) / 1e3:.2f} us")
    print("  per-CTA medians (us since CTA start): " + "; ".join(f"{n} {m / 1e3:.2f}" for n, m in zip(names_a, np.median(seg, axis=0))))
