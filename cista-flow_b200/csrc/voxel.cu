// Event stream -> voxel grid binning + normalisation (part 1 of the hot path).
//
// Replaces events_to_voxel_grid / _pytorch / _pol and event_preprocess(_pytorch)
// of the reference (utils/event_process.py:15-72, 127-190, 75-123, 193-239).
//
// Data layout in HBM: events float64 [total,4] rows (t,x,y,p), B windows
// concatenated + int64 offsets[B+1]; out float32 [B,nb,H,W] ([B,nb,2,H,W] for
// the polarity flavour).  One event is one 32-byte row = two 128-bit loads.
// Roofline: HBM.  Algorithmic bytes per window = 32*N_events + 4*nb*H*W.
//
// Two accumulation modes:
//   ATOMIC         scatter with fp32 RED.ADD resolved in L2 (the grid of a
//                  window is written once, stays L2-resident while its events
//                  stream in, and is normalised in place); sum order is
//                  unspecified -> agrees with the reference to <= 1e-5.
//   DETERMINISTIC  bit-exact.  The reference accumulates every cell
//                  sequentially in event order, all bin-ti ("left")
//                  contributions before all bin-ti+1 ("right") ones (two
//                  scatter-adds).  We reproduce exactly that order: a stable
//                  LSB radix sort of (cell-column key, event index) groups the
//                  events of each pixel in event order, then one thread per
//                  pixel walks its run bin by bin with un-contracted IEEE
//                  adds (__fadd_rn / fp64 add + round for the NumPy flavour).
// Time normalisation is fp64 with the reference's operation order
// (mul, then div) in both modes, so bin assignment is identical.
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
    const double tn = __ddiv_rn(__dmul_rn((double)(nb - 1), __dsub_rn(e.t, w.t0)), w.span);
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

// ------------------------------------------------------------ atomic mode ---
constexpr int kScatterThreads = 256;
constexpr int kScatterUnroll = 4;

__global__ void __launch_bounds__(kScatterThreads)
voxel_scatter_atomic_kernel(const double *__restrict__ ev, const int64_t *__restrict__ off, int64_t total,
                            int B, int nb, int H, int W, int flavour, float *__restrict__ out) {
    const int64_t tile = (int64_t)blockIdx.x * (kScatterThreads * kScatterUnroll);
    const int64_t plane = (int64_t)H * W;
    const int planes_per_bin = flavour == CF_FLAVOUR_POL ? 2 : 1;
    Window w;
    w.b = -1;
    Event e[kScatterUnroll];
#pragma unroll
    for (int k = 0; k < kScatterUnroll; ++k) {  // all loads in flight first
        const int64_t i = tile + k * kScatterThreads + threadIdx.x;
        if (i < total) e[k] = load_event(ev, i);
    }
#pragma unroll
    for (int k = 0; k < kScatterUnroll; ++k) {
        const int64_t i = tile + k * kScatterThreads + threadIdx.x;
        if (i >= total) break;
        locate_window(w, i, off, ev, B);
        const Binned b = bin_event(e[k], w, nb, H, W, flavour);
        if (!b.ok) continue;
        float wl, wr;
        if (flavour == CF_FLAVOUR_TORCH) {
            weights_f32(b, wl, wr);
        } else {
            double dl, dr;
            weights_f64(b, dl, dr);
            wl = (float)dl;
            wr = (float)dr;
        }
        float *cell = out + (((int64_t)w.b * nb + b.bin) * planes_per_bin + b.chan) * plane + (int64_t)b.y * W + b.x;
        atomicAdd(cell, wl);  // result unused -> RED.E.ADD.F32
        if (b.bin + 1 < nb) atomicAdd(cell + planes_per_bin * plane, wr);
    }
}

// ------------------------------------------------------ deterministic mode ---
constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096 keys per CTA
constexpr int kRsWarps = kRsThreads / 32;

// key = pixel column of the output ((b*planes + chan)*H*W + y*W + x); events
// that contribute nothing get key == invalid_key (sorted to the very end).
__global__ void __launch_bounds__(256)
det_make_keys_kernel(const double *__restrict__ ev, const int64_t *__restrict__ off, int64_t total,
                     int B, int nb, int H, int W, int flavour, uint32_t invalid_key,
                     uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    Window w;
    w.b = -1;
    locate_window(w, i, off, ev, B);
    const Binned b = bin_event(load_event(ev, i), w, nb, H, W, flavour);
    const int planes = flavour == CF_FLAVOUR_POL ? 2 : 1;
    uint32_t key = invalid_key;
    if (b.ok) key = (uint32_t)(((int64_t)w.b * planes + b.chan) * ((int64_t)H * W) + (int64_t)b.y * W + b.x);
    keys[i] = key;
    vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(kRsThreads)
rs_hist_kernel(const uint32_t *__restrict__ keys, int64_t n, int shift, uint32_t *__restrict__ hist, int nblk) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kRsTile;
#pragma unroll
    for (int k = 0; k < kRsItems; ++k) {
        const int64_t i = base + k * kRsThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);  // integer: order-independent
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `n` counters in place, single CTA of 1024 threads
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t *__restrict__ data, int64_t n) {
    __shared__ uint32_t warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t per = (n + 1023) / 1024;
    const int64_t s = (int64_t)tid * per, e = min(n, s + per);
    uint32_t sum = 0;
    for (int64_t i = s; i < e; ++i) sum += data[i];
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t t = warp_tot[lane], ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += v;
        }
        warp_tot[lane] = ti - t;
    }
    __syncthreads();
    uint32_t run = warp_tot[warp] + inc - sum;
    for (int64_t i = s; i < e; ++i) {
        const uint32_t v = data[i];
        data[i] = run;
        run += v;
    }
}

// Stable scatter of one 8-bit digit.  Order inside a CTA tile: warp-major,
// round-major, lane-minor == global index order; __match_any_sync ranks the
// lanes of a round, a per-warp running count ranks the rounds, a scan over the
// warps ranks the warps, the scanned histogram ranks the CTAs.
__global__ void __launch_bounds__(kRsThreads)
rs_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                  uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n, int shift,
                  const uint32_t *__restrict__ hist, int nblk) {
    __shared__ uint32_t warp_cnt[kRsWarps][256];
    __shared__ uint32_t gbase[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < kRsWarps * 256; k += kRsThreads) (&warp_cnt[0][0])[k] = 0;
    __syncthreads();
    const int64_t wbase = (int64_t)blockIdx.x * kRsTile + (int64_t)warp * (kRsItems * 32);
    uint32_t key[kRsItems], rnk[kRsItems];
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? keys_in[i] : 0u;
        const uint32_t d = valid ? ((key[r] >> shift) & 255u) : (256u + lane);  // invalid lanes never match
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t before = __popc(peers & ((1u << lane) - 1u));
        uint32_t prev = 0;
        if (valid) prev = warp_cnt[warp][d];
        __syncwarp();
        if (valid && before == 0) warp_cnt[warp][d] = prev + __popc(peers);
        __syncwarp();
        rnk[r] = prev + before;
    }
    __syncthreads();
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w) {
            const uint32_t c = warp_cnt[w][tid];
            warp_cnt[w][tid] = run;
            run += c;
        }
        gbase[tid] = hist[(size_t)tid * nblk + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (key[r] >> shift) & 255u;
            const uint32_t dst = gbase[d] + warp_cnt[warp][d] + rnk[r];
            keys_out[dst] = key[r];
            vals_out[dst] = vals_in[i];
        }
    }
}

// One thread per run of equal keys (= one output pixel column).  The run is in
// event order; bin indices are non-decreasing along it.  For each bin: first
// the left weights of the events with ti == bin, then the right weights of the
// events with ti == bin-1 -- the order of the reference's two scatter-adds.
__global__ void __launch_bounds__(256)
det_accumulate_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, int64_t n,
                      uint32_t invalid_key, const double *__restrict__ ev, const int64_t *__restrict__ off,
                      int B, int nb, int H, int W, int flavour, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = keys[i];
    if (k == invalid_key) return;
    if (i > 0 && keys[i - 1] == k) return;  // not the head of its run

    const int64_t plane = (int64_t)H * W;
    const int planes = flavour == CF_FLAVOUR_POL ? 2 : 1;
    const int64_t col = (int64_t)k / plane;  // b*planes + chan
    const int64_t pix = (int64_t)k - col * plane;
    const int b = (int)(col / planes), chan = (int)(col - (int64_t)b * planes);
    Window w;
    w.b = -1;
    locate_window(w, vals[i], off, ev, B);
    float *cell0 = out + ((int64_t)b * nb * planes + chan) * plane + pix;  // bin 0 of this column

    int64_t pos = i, prev_s = i, prev_e = i;
    for (int bin = 0; bin < nb; ++bin) {
        float acc = 0.f;
        bool touched = false;
        int64_t e = pos;
        while (e < n && keys[e] == k) {  // left weights, ti == bin
            const Binned bb = bin_event(load_event(ev, vals[e]), w, nb, H, W, flavour);
            if (bb.bin != bin) break;
            if (flavour == CF_FLAVOUR_TORCH) {
                float wl, wr;
                weights_f32(bb, wl, wr);
                acc = __fadd_rn(acc, wl);
            } else {
                double wl, wr;
                weights_f64(bb, wl, wr);
                acc = (float)__dadd_rn((double)acc, wl);
            }
            touched = true;
            ++e;
        }
        for (int64_t j = prev_s; j < prev_e; ++j) {  // right weights, ti == bin-1
            const Binned bb = bin_event(load_event(ev, vals[j]), w, nb, H, W, flavour);
            if (flavour == CF_FLAVOUR_TORCH) {
                float wl, wr;
                weights_f32(bb, wl, wr);
                acc = __fadd_rn(acc, wr);
            } else {
                double wl, wr;
                weights_f64(bb, wl, wr);
                acc = (float)__dadd_rn((double)acc, wr);
            }
            touched = true;
        }
        if (touched) cell0[(int64_t)bin * planes * plane] = acc;
        prev_s = pos;
        prev_e = e;
        pos = e;
        if (prev_s == prev_e && (pos >= n || keys[pos] != k)) break;
    }
}

// ---------------------------------------------- statistics + normalisation ---
struct alignas(16) Partial {
    double sum, sumsq;
    long long nnz;
    float mn, mx;
};

constexpr int kStatThreads = 256;
constexpr int kMaxChunks = 256;

__device__ __forceinline__ float hot_filter(float v, float thr) { return (thr > 0.f && fabsf(v) > thr) ? 0.f : v; }

__global__ void __launch_bounds__(kStatThreads)
voxel_stats_kernel(const float *__restrict__ grid, int64_t cells, int64_t chunk_len, float hot_thr,
                   Partial *__restrict__ partials, int chunks) {
    const int b = blockIdx.y, c = blockIdx.x;
    const float *g = grid + (int64_t)b * cells;
    const int64_t s = (int64_t)c * chunk_len, e = min(cells, s + chunk_len);
    double sum = 0.0, sumsq = 0.0;
    long long nnz = 0;
    float mn = INFINITY, mx = -INFINITY;
    auto take = [&](float raw) {
        const float v = hot_filter(raw, hot_thr);
        sum += (double)v;
        sumsq += (double)v * (double)v;
        nnz += (v != 0.f);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    };
    if (((cells | chunk_len) & 3) == 0) {  // 128-bit loads (chunk starts stay 16-byte aligned)
        const float4 *g4 = reinterpret_cast<const float4 *>(g);
        for (int64_t i = (s >> 2) + threadIdx.x; i < (e >> 2); i += kStatThreads) {
            const float4 q = __ldg(g4 + i);
            take(q.x); take(q.y); take(q.z); take(q.w);
        }
    } else {
        for (int64_t i = s + threadIdx.x; i < e; i += kStatThreads) take(__ldg(g + i));
    }
    // fixed-shape tree: deterministic for a given launch geometry
    sum = warp_sum(sum); sumsq = warp_sum(sumsq); nnz = warp_sum(nnz);
    mn = warp_min(mn); mx = warp_max(mx);
    __shared__ Partial sh[kStatThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = Partial{sum, sumsq, nnz, mn, mx};
    __syncthreads();
    if (threadIdx.x == 0) {
        Partial t = sh[0];
        for (int k = 1; k < kStatThreads / 32; ++k) {
            t.sum += sh[k].sum; t.sumsq += sh[k].sumsq; t.nnz += sh[k].nnz;
            t.mn = fminf(t.mn, sh[k].mn); t.mx = fmaxf(t.mx, sh[k].mx);
        }
        partials[(size_t)b * chunks + c] = t;
    }
}

__global__ void __launch_bounds__(kStatThreads)
voxel_normalise_kernel(const float *in, float *out /* may alias in */, int64_t cells, int64_t chunk_len,
                       float hot_thr, int mode, const Partial *__restrict__ partials, int chunks) {
    const int b = blockIdx.y, c = blockIdx.x;
    __shared__ double s_a, s_b;   // out = (v - s_a) * s_b
    __shared__ int s_identity;
    if (threadIdx.x < 32) {
        // every CTA of a window re-reduces the window's partials in the same fixed order
        const Partial *p = partials + (size_t)b * chunks;
        double sum = 0.0, sumsq = 0.0;
        long long nnz = 0;
        float mn = INFINITY, mx = -INFINITY;
        for (int k = threadIdx.x; k < chunks; k += 32) {
            sum += p[k].sum; sumsq += p[k].sumsq; nnz += p[k].nnz;
            mn = fminf(mn, p[k].mn); mx = fmaxf(mx, p[k].mx);
        }
        sum = warp_sum(sum); sumsq = warp_sum(sumsq); nnz = warp_sum(nnz);
        mn = warp_min(mn); mx = warp_max(mx);
        if (threadIdx.x == 0) {
            if (mode == CF_PRE_STD) {
                s_identity = nnz == 0;  // event_process.py:205 -- untouched when there is no non-zero entry
                const double mean = nnz ? sum / (double)nnz : 0.0;
                const double var = nnz ? sumsq / (double)nnz - mean * mean : 0.0;
                s_a = mean;
                s_b = 1.0 / (sqrt(fmax(var, 0.0)) + 1e-8);
            } else {
                s_identity = 0;
                s_a = (double)mn;
                s_b = 1.0 / ((double)mx - (double)mn + 1e-8);
            }
        }
    }
    __syncthreads();
    const double a = s_a, inv = s_b;
    const bool identity = s_identity != 0;
    const float *g = in + (int64_t)b * cells;
    float *o = out + (int64_t)b * cells;
    const int64_t s = (int64_t)c * chunk_len, e = min(cells, s + chunk_len);
    // one fp64 subtract + multiply per NON-ZERO cell (a true fp64 divide per cell made this
    // kernel 3x slower than the scatter itself); the result is rounded once to fp32
    auto norm = [&](float raw) -> float {
        const float v = hot_filter(raw, hot_thr);
        if (identity) return v;
        if (mode == CF_PRE_STD) return (v != 0.f) ? (float)(((double)v - a) * inv) : 0.f;
        return (float)(((double)v - a) * inv);
    };
    if (((cells | chunk_len) & 3) == 0) {
        const float4 *g4 = reinterpret_cast<const float4 *>(g);
        float4 *o4 = reinterpret_cast<float4 *>(o);
        for (int64_t i = (s >> 2) + threadIdx.x; i < (e >> 2); i += kStatThreads) {
            const float4 q = g4[i];
            o4[i] = make_float4(norm(q.x), norm(q.y), norm(q.z), norm(q.w));
        }
    } else {
        for (int64_t i = s + threadIdx.x; i < e; i += kStatThreads) o[i] = norm(g[i]);
    }
}

static void stat_geometry(int B, int64_t cells, int &chunks, int64_t &chunk_len) {
    // enough CTAs to fill 148 SMs twice, chunks of >= 4096 cells, <= kMaxChunks per window
    int64_t want = ceil_div(2 * 148, B > 0 ? B : 1);
    int64_t cap = ceil_div(cells, 4096);
    chunks = (int)(want < 1 ? 1 : want);
    if (chunks > cap) chunks = (int)(cap < 1 ? 1 : cap);
    if (chunks > kMaxChunks) chunks = kMaxChunks;
    chunk_len = ceil_div(cells, chunks);
    chunk_len = (chunk_len + 3) & ~(int64_t)3;  // keeps every chunk start 16-byte aligned
}

static int run_preprocess(const float *in, float *out, int B, int64_t cells, int preprocess, float hot_thr,
                          void *ws, size_t ws_bytes, cudaStream_t stream) {
    CF_REQUIRE(B <= 65535, CF_ERR_INVALID_ARG, "preprocess: B > 65535");
    CF_REQUIRE(ws && ws_bytes >= (size_t)B * kMaxChunks * sizeof(Partial), CF_ERR_WORKSPACE,
               "preprocess: workspace too small (%zu < %zu)", ws_bytes, (size_t)B * kMaxChunks * sizeof(Partial));
    CF_REQUIRE(aligned16(ws), CF_ERR_ALIGN, "preprocess: workspace not 16-byte aligned");
    int chunks;
    int64_t chunk_len;
    stat_geometry(B, cells, chunks, chunk_len);
    Partial *partials = reinterpret_cast<Partial *>(ws);
    dim3 grid(chunks, B);
    voxel_stats_kernel<<<grid, kStatThreads, 0, stream>>>(in, cells, chunk_len, hot_thr, partials, chunks);
    CF_LAUNCH_CHECK("voxel_stats_kernel");
    voxel_normalise_kernel<<<grid, kStatThreads, 0, stream>>>(in, out, cells, chunk_len, hot_thr, preprocess, partials, chunks);
    CF_LAUNCH_CHECK("voxel_normalise_kernel");
    return CF_OK;
}

static int radix_bits(uint64_t max_key) {
    int bits = 1;
    while (bits < 32 && (max_key >> bits) != 0) ++bits;
    return bits;
}

struct DetLayout {
    size_t keys0, vals0, keys1, vals1, hist, end;
    int nblk;
};
static DetLayout det_layout(int64_t total, size_t base) {
    DetLayout L;
    L.nblk = (int)ceil_div(total > 0 ? total : 1, kRsTile);
    const size_t arr = align_up((size_t)(total > 0 ? total : 1) * sizeof(uint32_t), 256);
    L.keys0 = base;
    L.vals0 = L.keys0 + arr;
    L.keys1 = L.vals0 + arr;
    L.vals1 = L.keys1 + arr;
    L.hist = L.vals1 + arr;
    L.end = L.hist + align_up((size_t)256 * L.nblk * sizeof(uint32_t), 256);
    return L;
}

}  // namespace cf

extern "C" size_t cf_preprocess_workspace_bytes(int B, int64_t) {
    return (size_t)(B > 0 ? B : 1) * cf::kMaxChunks * sizeof(cf::Partial);
}

extern "C" size_t cf_voxel_workspace_bytes(int64_t total_events, int B, int, int, int, int mode, int, int) {
    size_t base = cf::align_up(cf_preprocess_workspace_bytes(B, 0), 256);
    if (mode == CF_VOXEL_DETERMINISTIC) return cf::det_layout(total_events, base).end;
    return base;
}

extern "C" int cf_voxel_preprocess(const float *in, float *out, int B, int64_t cells, int preprocess, float hot_thr,
                                   void *ws, size_t ws_bytes, cf_stream_t stream_) {
    using namespace cf;
    if (int rc = check_device()) return rc;
    CF_REQUIRE(in && out, CF_ERR_NULL, "cf_voxel_preprocess: null pointer");
    CF_REQUIRE(B >= 0 && cells > 0, CF_ERR_INVALID_ARG, "cf_voxel_preprocess: bad shape");
    CF_REQUIRE(preprocess == CF_PRE_STD || preprocess == CF_PRE_MAXMIN, CF_ERR_INVALID_ARG,
               "cf_voxel_preprocess: mode must be CF_PRE_STD or CF_PRE_MAXMIN");
    if (B == 0) return CF_OK;
    return run_preprocess(in, out, B, cells, preprocess, hot_thr, ws, ws_bytes, (cudaStream_t)stream_);
}

extern "C" int cf_voxel_bin(const double *events, const int64_t *offsets, int64_t total, int B, int nb, int H, int W,
                            int mode, int flavour, int preprocess, float hot_thr, float *out,
                            void *ws, size_t ws_bytes, cf_stream_t stream_) {
    using namespace cf;
    if (int rc = check_device()) return rc;
    CF_REQUIRE(out && offsets, CF_ERR_NULL, "cf_voxel_bin: null pointer");
    CF_REQUIRE(total == 0 || events, CF_ERR_NULL, "cf_voxel_bin: events is null");
    CF_REQUIRE(nb > 0 && H > 0 && W > 0, CF_ERR_INVALID_ARG, "cf_voxel_bin: num_bins, width, height must be > 0");
    CF_REQUIRE(B >= 0 && total >= 0, CF_ERR_INVALID_ARG, "cf_voxel_bin: negative size");
    CF_REQUIRE(mode == CF_VOXEL_ATOMIC || mode == CF_VOXEL_DETERMINISTIC, CF_ERR_INVALID_ARG, "cf_voxel_bin: bad mode %d", mode);
    CF_REQUIRE(flavour >= CF_FLAVOUR_TORCH && flavour <= CF_FLAVOUR_POL, CF_ERR_INVALID_ARG, "cf_voxel_bin: bad flavour %d", flavour);
    CF_REQUIRE(preprocess >= CF_PRE_NONE && preprocess <= CF_PRE_MAXMIN, CF_ERR_INVALID_ARG, "cf_voxel_bin: bad preprocess %d", preprocess);
    CF_REQUIRE(total == 0 || aligned16(events), CF_ERR_ALIGN, "cf_voxel_bin: events not 16-byte aligned");
    CF_REQUIRE(total < (1ll << 32), CF_ERR_INVALID_ARG, "cf_voxel_bin: more than 2^32 events in one call");
    if (B == 0) return CF_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int planes = flavour == CF_FLAVOUR_POL ? 2 : 1;
    const int64_t cells = (int64_t)nb * planes * H * W;
    CF_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * cells, stream));

    if (total > 0 && mode == CF_VOXEL_ATOMIC) {
        const int64_t blocks = ceil_div(total, kScatterThreads * kScatterUnroll);
        voxel_scatter_atomic_kernel<<<(unsigned)blocks, kScatterThreads, 0, stream>>>(events, offsets, total, B, nb, H, W, flavour, out);
        CF_LAUNCH_CHECK("voxel_scatter_atomic_kernel");
    } else if (total > 0) {
        const uint64_t columns = (uint64_t)B * planes * H * W;  // invalid key == columns
        CF_REQUIRE(columns < 0xffffffffull, CF_ERR_INVALID_ARG, "cf_voxel_bin: B*H*W too large for the deterministic mode");
        const size_t base = align_up(cf_preprocess_workspace_bytes(B, 0), 256);
        const DetLayout L = det_layout(total, base);
        CF_REQUIRE(ws && ws_bytes >= L.end, CF_ERR_WORKSPACE, "cf_voxel_bin: workspace too small (%zu < %zu)", ws_bytes, L.end);
        CF_REQUIRE(aligned16(ws), CF_ERR_ALIGN, "cf_voxel_bin: workspace not 16-byte aligned");
        char *w8 = reinterpret_cast<char *>(ws);
        uint32_t *k0 = (uint32_t *)(w8 + L.keys0), *v0 = (uint32_t *)(w8 + L.vals0);
        uint32_t *k1 = (uint32_t *)(w8 + L.keys1), *v1 = (uint32_t *)(w8 + L.vals1);
        uint32_t *hist = (uint32_t *)(w8 + L.hist);
        const unsigned eb = (unsigned)ceil_div(total, 256);
        det_make_keys_kernel<<<eb, 256, 0, stream>>>(events, offsets, total, B, nb, H, W, flavour, (uint32_t)columns, k0, v0);
        CF_LAUNCH_CHECK("det_make_keys_kernel");
        const int passes = (radix_bits(columns) + 7) / 8;
        for (int p = 0; p < passes; ++p) {
            rs_hist_kernel<<<L.nblk, kRsThreads, 0, stream>>>(k0, total, 8 * p, hist, L.nblk);
            CF_LAUNCH_CHECK("rs_hist_kernel");
            rs_scan_kernel<<<1, 1024, 0, stream>>>(hist, (int64_t)256 * L.nblk);
            CF_LAUNCH_CHECK("rs_scan_kernel");
            rs_scatter_kernel<<<L.nblk, kRsThreads, 0, stream>>>(k0, v0, k1, v1, total, 8 * p, hist, L.nblk);
            CF_LAUNCH_CHECK("rs_scatter_kernel");
            uint32_t *t = k0; k0 = k1; k1 = t;
            t = v0; v0 = v1; v1 = t;
        }
        det_accumulate_kernel<<<eb, 256, 0, stream>>>(k0, v0, total, (uint32_t)columns, events, offsets, B, nb, H, W, flavour, out);
        CF_LAUNCH_CHECK("det_accumulate_kernel");
    }
    if (preprocess != CF_PRE_NONE)
        return run_preprocess(out, out, B, cells, preprocess, hot_thr, ws, ws_bytes, stream);
    return CF_OK;
}
