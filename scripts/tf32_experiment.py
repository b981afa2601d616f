"""GPU experiment: error statistics of the tcgen05 TF32 correlation under the
debug flags of corr_build_tc.cu (run once per flag value: the env var is read once).
  CF_TC_FLAGS bit0: skip the round-to-nearest pre-pass (hardware truncation)
              bit1: TFLOAT32 tensor-map data type
              bit2: do not fuse the level-1 pooling
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cistaflow_b200 as cf
from cistaflow_b200 import synth

dev = torch.device("cuda", 0)
for (h_img, w_img, batch) in ((180, 240, 2), (260, 346, 1)):
    f1, f2, _ = synth.corr_inputs(batch, h_img, w_img, seed=9)
    a, b = torch.from_numpy(f1).to(dev), torch.from_numpy(f2).to(dev)
    B, D, h, w = a.shape
    N = h * w
    exact = torch.einsum("bdi,bdj->bij", a.reshape(B, D, N).double(), b.reshape(B, D, N).double()) / 16.0
    try:
        pyr = cf.build_pyramid(a, b, 4, precision="tf32")
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"flags={os.environ.get('CF_TC_FLAGS', '0')} {h}x{w}: FAILED {type(e).__name__}: {e}")
        sys.exit(1)
    got = pyr[0].view(B, N, N).double()
    err = got - exact
    rel_signed = (err * torch.sign(exact)).mean().item() / exact.abs().mean().item()
    l1_ref = torch.nn.functional.avg_pool2d(pyr[0], 2, 2)
    print(f"flags={os.environ.get('CF_TC_FLAGS', '0')} {h}x{w} B={B}: max|err|={err.abs().max().item():.3e} "
          f"max|ref|={exact.abs().max().item():.3f} normwise={err.abs().max().item() / exact.abs().max().item():.3e} "
          f"rms_err={err.pow(2).mean().sqrt().item():.3e} signed_bias_rel={rel_signed:.3e} "
          f"l1_vs_pool={(pyr[1] - l1_ref).abs().max().item():.2e}")
