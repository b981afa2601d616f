"""fp16-operand correlation (CF_CORR_F16) against the fp32 SIMT kernel, and its timing next to TF32.
    python scripts/f16_check.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cistaflow_b200 as cf  # noqa: E402
from cistaflow_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for (H, W, B, scale) in ((192, 256, 2, 1.0), (288, 352, 2, 1.0), (480, 640, 1, 1.0), (192, 256, 2, 3.0e6), (192, 256, 2, 1.0e-9)):
    f1, f2, _ = synth.corr_inputs(B, H, W, 5)
    a, b = torch.from_numpy(f1).to(dev) * scale, torch.from_numpy(f2).to(dev)
    ref = cf.build_pyramid(a, b, 4, precision="fp32")
    for prec in ("tf32", "f16"):
        got = cf.build_pyramid(a, b, 4, precision=prec)
        errs = [((g - r).abs().max() / r.abs().max()).item() for g, r in zip(got, ref)]
        print(f"{H}x{W} B={B} x{scale:g} {prec:5s} max|err|/max|ref| per level:", " ".join(f"{e:.2e}" for e in errs))
for (H, W, B) in ((480, 640, 8), (180, 240, 64), (624, 970, 1), (260, 346, 64)):
    f1, f2, _ = synth.corr_inputs(B, H, W, 1)
    a, b = torch.from_numpy(f1).to(dev), torch.from_numpy(f2).to(dev)
    h, w = a.shape[2], a.shape[3]
    out = [torch.empty((B * h * w, 1, h >> l, w >> l), device=dev) for l in range(4)]
    t = {p: timeit(lambda p=p: cf.build_pyramid(a, b, 4, precision=p, out=out)) for p in ("tf32", "f16", "auto")}
    print(f"corr_build {H}x{W} B={B}: " + "  ".join(f"{p} {v:7.1f} us" for p, v in t.items()))
