import os, sys, torch
sys.path.insert(0, "/root/repo")
import cistaflow_b200 as cf
from cistaflow_b200 import synth
dev = torch.device("cuda", 0)
def timeit(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(4): fn()
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(reps): g.replay()
        b.record(s)
        torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * 4) * 1e3
for (H, W, B) in ((180, 240, 8), (480, 640, 8)):
    sets = [tuple(torch.from_numpy(a).to(dev) for a in synth.warp_inputs(B, H, W, 3 + k, 128, flow_kind="smooth")) for k in range(4)]
    i = [0]
    def pick():
        i[0] = (i[0] + 1) % 4
        return sets[i[0]]
    ds = [torch.nn.functional.interpolate(f, scale_factor=0.5, mode="bilinear", align_corners=True) for (_, _, f) in sets]
    t_all = timeit(lambda: cf.warp_frame_and_codes(*pick(), "forward"))
    def codes_only():
        img, codes, flow = pick()
        cf.warp(codes, flow, -1.0)
    def img_only():
        img, codes, flow = pick()
        cf.warp(img, flow, -1.0)
    print(f"{H}x{W} B={B}: fused {t_all:.1f} us, codes only {timeit(codes_only):.1f} us, image only {timeit(img_only):.1f} us")
