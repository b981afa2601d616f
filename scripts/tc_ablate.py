"""Ablation timings of the tcgen05 correlation kernel (CF_TC_FLAGS bits 8-10, experiments only) plus two
reference points measured the same way: a pure HBM write stream and cuBLAS TF32.
    python scripts/tc_ablate.py            (run once per CF_TC_FLAGS value; prints us per launch)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cistaflow_b200 as cf
from cistaflow_b200 import synth

dev = torch.device("cuda", 0)


def timeit(fn, reps=20):
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


print("CF_TC_FLAGS =", os.environ.get("CF_TC_FLAGS", "0"))
for (H, W, B) in ((480, 640, 8), (180, 240, 64), (624, 970, 1)):
    f1, f2, _ = synth.corr_inputs(B, H, W, seed=1)
    a, b = torch.from_numpy(f1).to(dev), torch.from_numpy(f2).to(dev)
    print(f"  corr_build {H}x{W} B={B}: {timeit(lambda: cf.build_pyramid(a, b, 4)):8.1f} us")
if os.environ.get("CF_TC_FLAGS", "0") == "0":
    x = torch.empty(1 << 30, dtype=torch.float32, device=dev)  # 4 GB
    us = timeit(lambda: x.fill_(1.0), 5)
    print(f"  fill_ 4 GB: {us:8.1f} us = {x.numel() * 4 / us / 1e3:7.1f} GB/s (pure write stream)")
    y = torch.empty(1 << 28, dtype=torch.float32, device=dev)
    z = torch.empty_like(y)
    us = timeit(lambda: z.copy_(y), 5)
    print(f"  copy 1+1 GB: {us:8.1f} us = {2 * y.numel() * 4 / us / 1e3:7.1f} GB/s (read + write)")
    del x, y, z
    torch.backends.cuda.matmul.allow_tf32 = True
    m = torch.randn(8192, 8192, device=dev)
    n = torch.randn(8192, 8192, device=dev)
    us = timeit(lambda: m @ n, 10)
    print(f"  cuBLAS tf32 8192^3: {us:8.1f} us = {2 * 8192 ** 3 / us / 1e6:7.1f} TFLOP/s")
