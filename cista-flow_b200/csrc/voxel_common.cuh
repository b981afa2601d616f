// Event decoding, time binning and window statistics shared by voxel.cu (L2-atomic and deterministic
// paths) and voxel_tiled.cu (partition + shared-memory accumulation).  The arithmetic follows
// utils/event_process.py:39-66 (NumPy), :152-187 (torch), :193-239 (preprocess) operation by operation.
#pragma once

#include "common.cuh"

namespace cf {

// ------------------------------------------------------------------ events ---
struct Event {
    double t, x, y, p;
};

__device__ __forceinline__ Event load_event(const double *__restrict__ ev, int64_t i) {
    const double2 *p = reinterpret_cast<const double2 *>(ev) + 2 * i;
    const double2 a = __ldg(p), b = __ldg(p + 1);  // 2 x 128-bit
    return Event{a.x, a.y, b.x, b.y};
}

struct Window {
    int b;
    int64_t begin, end;
    double t0, span;
};

// Window that owns event i, starting the search from hint `w.b`.
__device__ __forceinline__ void locate_window(Window &w, int64_t i, const int64_t *__restrict__ off,
                                              const double *__restrict__ ev, int B) {
    if (w.b >= 0 && i >= w.begin && i < w.end) return;
    int lo = 0, hi = B - 1;  // last b with off[b] <= i
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(off + mid) <= i) lo = mid; else hi = mid - 1;
    }
    w.b = lo;
    w.begin = __ldg(off + lo);
    w.end = __ldg(off + lo + 1);
    w.t0 = __ldg(ev + 4 * w.begin);
    const double last = __ldg(ev + 4 * (w.end - 1));
    w.span = __dsub_rn(last, w.t0);
    if (w.span == 0.0) w.span = 1.0;  // event_process.py:43-44
}

struct Binned {
    int bin;      // ti
    int x, y;     // truncated coordinates
    int chan;     // polarity channel (POL flavour)
    double dt;    // t* - ti  (fp64)
    double sgn;   // +-1 weight sign (p itself when p != 0)
    bool ok;
};

__device__ __forceinline__ Binned bin_event(const Event &e, const Window &w, int nb, int H, int W, int flavour) {
    Binned r;
    // t* = (nb-1)*(t-t0)/dT : one rounding per operation, no contraction
    // (MVSEC flavour: ((t-t0)/dT)*(nb-1), MVSEC_utils.py:281-282)
    const double tn = flavour == CF_FLAVOUR_MVSEC
                          ? __dmul_rn(__ddiv_rn(__dsub_rn(e.t, w.t0), w.span), (double)(nb - 1))
                          : __ddiv_rn(__dmul_rn((double)(nb - 1), __dsub_rn(e.t, w.t0)), w.span);
    const double lo = floor(tn);
    r.ok = (lo >= 0.0) && (lo < (double)nb) && (e.x >= 0.0) && (e.y >= 0.0) && (e.x < (double)W) && (e.y < (double)H);
    r.bin = r.ok ? (int)lo : 0;
    r.x = r.ok ? (int)e.x : 0;
    r.y = r.ok ? (int)e.y : 0;
    r.dt = __dsub_rn(tn, lo);
    r.chan = 0;
    if (flavour == CF_FLAVOUR_POL) {
        r.chan = (int)e.p;
        r.ok = r.ok && (e.p >= 0.0) && (e.p < 2.0);
        r.sgn = (e.p == 0.0) ? 1.0 : e.p;
    } else if (flavour == CF_FLAVOUR_MVSEC) {
        r.sgn = e.p;  // polarity / weight as given (MVSEC_utils.py:287): 0 contributes nothing
    } else {
        r.sgn = (e.p == 0.0) ? -1.0 : e.p;
    }
    return r;
}

// left/right weights exactly as the reference forms them
__device__ __forceinline__ void weights_f32(const Binned &b, float &wl, float &wr) {
    const float s = (float)b.sgn, f = (float)b.dt;  // event_process.py:163,169-170
    wl = __fmul_rn(s, __fsub_rn(1.0f, f));
    wr = __fmul_rn(s, f);
}
__device__ __forceinline__ void weights_f64(const Binned &b, double &wl, double &wr) {
    wl = __dmul_rn(b.sgn, __dsub_rn(1.0, b.dt));  // event_process.py:58-59
    wr = __dmul_rn(b.sgn, b.dt);
}

// MVSEC flavour: p * max(0, 1 - |t* - bin|) for bin = ti and ti + 1 (MVSEC_utils.py:286-287).  |t* - ti| = dt and
// (ti + 1) - t* = 1 - dt are both exact in fp64, so the two weights are p * fl(1 - dt) and p * fl(1 - fl(1 - dt)).
__device__ __forceinline__ void weights_mvsec(const Binned &b, double &wl, double &wr) {
    const double one_minus = __dsub_rn(1.0, b.dt);
    wl = __dmul_rn(b.sgn, one_minus);
    wr = __dmul_rn(b.sgn, __dsub_rn(1.0, one_minus));
}
// fp64 weights of the non-TORCH flavours
__device__ __forceinline__ void weights_wide(const Binned &b, int flavour, double &wl, double &wr) {
    if (flavour == CF_FLAVOUR_MVSEC) weights_mvsec(b, wl, wr);
    else weights_f64(b, wl, wr);
}

struct alignas(16) Partial {
    double sum, sumsq;
    long long nnz;
    float mn, mx;
};

__device__ __forceinline__ float hot_filter(float v, float thr) { return (thr > 0.f && fabsf(v) > thr) ? 0.f : v; }

constexpr int kMaxChunks = 512;  // partial-statistics slots per window in the workspace

}  // namespace cf
