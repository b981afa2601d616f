import os, sys, torch
sys.path.insert(0, "/root/repo")
import cistaflow_b200 as cf
from cistaflow_b200 import synth
dev = torch.device("cuda", 0)
H, W, B = (int(v) for v in sys.argv[1:4]); prec = sys.argv[4]
f1, f2, _ = synth.corr_inputs(B, H, W, 1)
a, b = torch.from_numpy(f1).to(dev), torch.from_numpy(f2).to(dev)
for _ in range(3):
    cf.build_pyramid(a, b, 4, precision=prec)
torch.cuda.synchronize()
