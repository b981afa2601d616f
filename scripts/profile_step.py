"""One hot-path step of the headline workload (BASELINE.json configs[4], this rank's share) through the public API, for ncu:

    python scripts/profile_step.py [--batch 64] [--steps 2]

Step k's launches are the same every step (printed: launches per step), so `ncu -s <launches per step> -c <launches per
step>` captures exactly the second step.  No CUDA graphs here (ncu serialises launches anyway)."""
import argparse
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--config", default="configs[4]")
args = ap.parse_args()
from cistaflow_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = bench.make_cfg(args.config)
hp = bench.HotPath(cfg, args.batch, dev, 1234 + 4000, n_sets=1)
lib = _lib.load()
for k in range(args.steps):
    n0 = lib.cf_launch_count()
    out = hp.step(hp.sets[0])
    torch.cuda.synchronize()
    print(f"step {k}: {lib.cf_launch_count() - n0} library launches", flush=True)
    del out
