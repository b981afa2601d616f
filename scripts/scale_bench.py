"""Per-kernel roofline fractions at every BASELINE.json config shape (not only configs[1]).

bench.py times the step at configs[1], where every kernel runs for 7-30 us and launch ramp/tail and
dependent-latency chains, not bandwidth, decide the fraction.  This script times each of the four
kernels alone at the shapes of configs[1..5]: algorithmic bytes / flops per
launch (same model as bench.py) over the CUDA-event duration.  Inputs rotate over enough sets to exceed
the 126 MB L2, launches are replayed from a CUDA graph (no host gaps).

    python scripts/scale_bench.py [--out gpurun_out/scale_bench.json] [--only warp,voxel,...] [--cases cfg5_b8,...]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cistaflow_b200 as cf  # noqa: E402
from cistaflow_b200 import synth  # noqa: E402

L2_BYTES = 126e6
CASES = {
    # name: (H, W, batch, events per window)
    "cfg2_180x240_b8": (180, 240, 8, 15000),
    "cfg2_180x240_b64": (180, 240, 64, 15000),
    "cfg3_260x346_b1": (260, 346, 1, 50000),
    "cfg3_260x346_b64": (260, 346, 64, 50000),
    "cfg5_480x640_b8": (480, 640, 8, 100000),
    "cfg5_480x640_b64": (480, 640, 64, 100000),
    "cfg4_624x970_b1": (624, 970, 1, 1000000),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"] / 2.0, "measured"
    return 6650.0, 1590.0 / 2.0, "fallback"


def graph_time(fns, stream, inner, reps=5):
    """fns: one closure per rotating input set.  Returns seconds per launch."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        for i in range(inner):
            fns[i % len(fns)]()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        g.replay()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * inner) * 1e-3


def n_sets(bytes_per_set, cap=6):
    return int(max(2, min(cap, np.ceil(2.2 * L2_BYTES / max(bytes_per_set, 1)))))


def tile_windows(ev, off, B):
    """B windows from a few distinct synthetic ones (keeps host generation cheap)."""
    base = len(off) - 1
    k = -(-B // base)
    evs = np.concatenate([ev] * k)
    offs = np.concatenate([[0], np.cumsum(np.tile(np.diff(off), k))]).astype(np.int64)[:B + 1]
    return evs[:offs[-1]], offs


def _maybe_set_l2_fetch_granularity():
    """Experiment knob: CF_L2_FETCH=32|64|128 sets cudaLimitMaxL2FetchGranularity (0x05) for the process."""
    g = os.environ.get("CF_L2_FETCH")
    if g:
        import ctypes
        torch.zeros(1, device="cuda")
        rt = ctypes.CDLL("libcudart.so.12")
        err = rt.cudaDeviceSetLimit(ctypes.c_int(5), ctypes.c_size_t(int(g)))
        val = ctypes.c_size_t(0)
        rt.cudaDeviceGetLimit(ctypes.byref(val), ctypes.c_int(5))
        print(f"cudaLimitMaxL2FetchGranularity <- {g}: err {err}, now {val.value}")


def run_cases(cases, only, dev=None, verbose=True, voxel_paths=True):
    """Times the kernels named in `only` at the shapes named in `cases`; returns (rows, peaks)."""
    if dev is None:
        dev = torch.device("cuda", 0)
        torch.cuda.set_device(dev)
    stream = torch.cuda.Stream(dev)
    hbm, tf32, src = peaks()
    rows = []
    with torch.cuda.stream(stream):
        for name in cases:
            H, W, B, nev = CASES[name]
            hp, wp = synth.padded_dims(H, W)
            h, w = hp // 8, wp // 8
            N = h * w
            lvl_cells = sum((h >> l) * (w >> l) for l in range(4))
            row = {"case": name, "H": H, "W": W, "batch": B, "events": nev, "N": N}

            if "voxel" in only:
                nbytes = B * (32 * nev + 4 * 5 * H * W)
                ns = n_sets(nbytes)
                sets = []
                for s in range(ns):
                    ev, off = synth.event_windows(min(B, 4), nev, H, W, seed=11 + s)
                    ev, off = tile_windows(ev, off, B)
                    sets.append((torch.from_numpy(ev).to(dev), torch.from_numpy(off).to(dev),
                                 torch.empty((B, 5, H, W), device=dev)))
                variants = (("voxel+norm", "std", "atomic"), ("voxel+norm[l2]", "std", "atomic_l2"),
                            ("voxel+norm[tiled]", "std", "atomic_tiled"), ("voxel_only", None, "atomic"))
                for label, norm, path in (variants if voxel_paths else variants[:1]):
                    fns = [(lambda e=e, o=o, out=out: cf.events_to_voxel_grid_batched(
                        e, o, 5, W, H, normalize=norm, filter_hot_pixel=norm is not None, flavour="numpy",
                        mode=path, out=out)) for (e, o, out) in sets]
                    t = graph_time(fns, stream, inner=2 * ns)
                    row[label] = {"us": t * 1e6, "GB/s": nbytes / t / 1e9, "frac_hbm": nbytes / t / 1e9 / hbm,
                                  "Mev/s": B * nev / t / 1e6, "bytes": nbytes}
                del sets

            if "warp" in only:
                nbytes = B * (8 * H * W + 8 * H * W + (H // 2) * (W // 2) * 8 * 128)
                ns = n_sets(nbytes, cap=4)
                sets = []
                for s in range(ns):
                    img, codes, flow = synth.warp_inputs(min(B, 2), H, W, 21 + s, 128, flow_kind="smooth")
                    rep = -(-B // img.shape[0])
                    t_ = [torch.from_numpy(np.concatenate([a] * rep)[:B]).to(dev) for a in (img, codes, flow)]
                    sets.append((*t_, torch.empty_like(t_[0]), torch.empty_like(t_[1])))
                fns = [(lambda i=i, z=z, f=f, oi=oi, oz=oz: cf.warp_frame_and_codes(i, z, f, "forward", out=(oi, oz)))
                       for (i, z, f, oi, oz) in sets]
                t = graph_time(fns, stream, inner=2 * ns)
                row["warp_frame_and_codes"] = {"us": t * 1e6, "GB/s": nbytes / t / 1e9, "frac_hbm": nbytes / t / 1e9 / hbm,
                                               "bytes": nbytes}
                del sets

            if "build" in only or "lookup" in only:
                bbytes = 4 * B * (2 * 256 * N + N * lvl_cells)
                flops = 2 * B * N * N * 256
                ns = n_sets(bbytes, cap=3)
                sets = []
                for s in range(ns):
                    f1, f2, c0 = synth.corr_inputs(min(B, 2), H, W, 31 + s)
                    rep = -(-B // f1.shape[0])
                    f1, f2, c0 = [torch.from_numpy(np.concatenate([a] * rep)[:B]).to(dev) for a in (f1, f2, c0)]
                    pyr = [torch.empty((B * N, 1, h >> l, w >> l), device=dev) for l in range(4)]
                    sets.append((f1, f2, c0, pyr))
                if "build" in only:
                    fns = [(lambda f1=f1, f2=f2, pyr=pyr: cf.build_pyramid(f1, f2, 4, out=pyr)) for (f1, f2, c0, pyr) in sets]
                    t = graph_time(fns, stream, inner=2 * ns)
                    row["corr_build"] = {"us": t * 1e6, "GB/s": bbytes / t / 1e9, "frac_hbm": bbytes / t / 1e9 / hbm,
                                         "TF/s_tf32": flops / t / 1e12, "frac_tf32": flops / t / 1e12 / tf32,
                                         "bytes": bbytes, "flops": flops}
                if "lookup" in only:
                    for (f1, f2, c0, pyr) in sets:
                        cf.build_pyramid(f1, f2, 4, out=pyr)
                    lbytes = B * N * 2904
                    outs = [torch.empty((B, 324, h, w), device=dev) for _ in sets]
                    fns = [(lambda pyr=pyr, c0=c0, o=o: cf.corr_lookup(pyr, c0, 4, out=o))
                           for (f1, f2, c0, pyr), o in zip(sets, outs)]
                    t = graph_time(fns, stream, inner=4 * ns)
                    row["corr_lookup"] = {"us": t * 1e6, "GB/s": lbytes / t / 1e9, "frac_hbm": lbytes / t / 1e9 / hbm,
                                          "bytes": lbytes}
                del sets
            torch.cuda.empty_cache()
            rows.append(row)
            if verbose:
                print(json.dumps(row), flush=True)
    return rows, {"hbm_gbs": hbm, "tf32_tflops": tf32, "source": src}


def main():
    _maybe_set_l2_fetch_granularity()
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "scale_bench.json"))
    ap.add_argument("--only", default="voxel,warp,build,lookup")
    ap.add_argument("--cases", default=",".join(CASES))
    args = ap.parse_args()
    rows, pk = run_cases(args.cases.split(","), set(args.only.split(",")))
    hbm, tf32, src = pk["hbm_gbs"], pk["tf32_tflops"], pk["source"]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"peaks": {"hbm_gbs": hbm, "tf32_tflops": tf32, "source": src}, "rows": rows}, open(args.out, "w"), indent=1)
    print(f"\n{'case':20s} {'kernel':22s} {'us':>10s} {'GB/s':>9s} {'%HBM':>6s}  extra")
    for r in rows:
        for k, v in r.items():
            if isinstance(v, dict):
                extra = f"{v['TF/s_tf32']:.0f} TF/s = {100 * v['frac_tf32']:.1f}% tf32" if "TF/s_tf32" in v else \
                    (f"{v['Mev/s']:.0f} Mev/s" if "Mev/s" in v else "")
                print(f"{r['case']:20s} {k:22s} {v['us']:10.1f} {v['GB/s']:9.0f} {100 * v['frac_hbm']:6.1f}  {extra}")


if __name__ == "__main__":
    main()
