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
//   ATOMIC         sum order unspecified -> agrees with the reference to <= 1e-5.
//                  fp32 RED.ADD resolved in L2; the batch is processed in chunks
//                  of windows whose grids + events fit in L2 (<= 48 MB), each chunk
//                  running zero -> scatter -> statistics -> normalise-in-place
//                  back to back, so a grid is read and written in L2 and reaches
//                  HBM exactly once.
//                  Measured dead end, kept behind CF_VOXEL_FLAGS=2 for the record:
//                  privatising the grid in the distributed shared memory of a
//                  thread-block cluster and scattering with
//                  red.shared::cluster.add.f32 (voxel_cluster_kernel).  DSMEM
//                  atomics sustain ~19 G adds/s chip-wide vs ~69 G/s for L2
//                  atomics on a B200, so that path is 2-4x SLOWER than the L2 one
//                  (scripts/voxel_microbench.py).
//   DETERMINISTIC  bit-exact.  The reference accumulates every cell
//                  sequentially in event order, all bin-ti ("left")
//                  contributions before all bin-ti+1 ("right") ones (two
//                  scatter-adds).  We reproduce exactly that order: a stable
//                  LSB radix sort of (cell-column key, event index) groups the
//                  events of each pixel in event order, then one thread per
//                  pixel walks its run bin by bin with un-contracted IEEE
//                  adds (__fadd_rn / fp64 add + round for the NumPy flavour).
// Measured and dropped (round 1): taking the std statistics out of the scatter itself -- returning atomics
// (ATOMG instead of RED) give each add's previous value o, and sum f(o + w) - f(o) telescopes to the final
// grid's sum / sum of squares / non-zero count, so the separate statistics pass disappears.  Correct (all parity
// tests green) but 1.4-4x SLOWER end to end on the B200 (8x180x240: 24.4 against 16.6 us; 64x260x346: 543 against
// 132 us): returning atomics cost far more than the pass over the L2-resident grid they save.
// Also measured and dropped: statistics + normalisation as ONE pass (a CTA keeps its <= 32 KB slice of a window in
// registers, publishes its partial, waits on a per-window arrival counter -- tickets make the wait deadlock-free --
// and normalises out of registers: one read, one write, one launch).  Parity-green but slower than the two
// kernels at every shape (8x180x240: 18.7 against 16.6 us; 8x480x640: 55.2 against 43.6 us): the ticket atomic
// sits in front of every load address and the waiting CTAs hold their SM slots idle.
// Time normalisation is fp64 with the reference's operation order
// (mul, then div) in both modes, so bin assignment is identical.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "voxel_common.cuh"

namespace cg = cooperative_groups;

namespace cf {

// ------------------------------------------------------------ atomic mode ---
constexpr int kScatterThreads = 256;
constexpr int kScatterUnroll = 4;

__device__ __forceinline__ void scatter_body(const double *__restrict__ ev, const int64_t *__restrict__ off,
                                             int B, int nb, int H, int W, int flavour, float *__restrict__ out, int b_first,
                                             int b_end, int blk, int nblk) {
    // the events of windows [b_first, b_end) (one L2-sized chunk of the batch), grid-stride over
    // tiles of 1024 events: the host does not know the chunk's event count (offsets are on the device)
    const int64_t ev_first = __ldg(off + b_first), ev_end = __ldg(off + b_end);
    const int64_t plane = (int64_t)H * W;
    const int planes_per_bin = flavour == CF_FLAVOUR_POL ? 2 : 1;
    constexpr int64_t kTileEvents = kScatterThreads * kScatterUnroll;
    Window w;
    w.b = -1;
    for (int64_t tile = ev_first + (int64_t)blk * kTileEvents; tile < ev_end; tile += (int64_t)nblk * kTileEvents) {
        Event e[kScatterUnroll];
#pragma unroll
        for (int k = 0; k < kScatterUnroll; ++k) {  // all loads in flight first
            const int64_t i = tile + k * kScatterThreads + threadIdx.x;
            if (i < ev_end) e[k] = load_event(ev, i);
        }
#pragma unroll
        for (int k = 0; k < kScatterUnroll; ++k) {
            const int64_t i = tile + k * kScatterThreads + threadIdx.x;
            if (i >= ev_end) break;
            locate_window(w, i, off, ev, B);
            const Binned b = bin_event(e[k], w, nb, H, W, flavour);
            if (!b.ok) continue;
            float wl, wr;
            if (flavour == CF_FLAVOUR_TORCH) {
                weights_f32(b, wl, wr);
            } else {
                double dl, dr;
                weights_wide(b, flavour, dl, dr);
                wl = (float)dl;
                wr = (float)dr;
            }
            float *cell = out + (((int64_t)w.b * nb + b.bin) * planes_per_bin + b.chan) * plane + (int64_t)b.y * W + b.x;
            atomicAdd(cell, wl);  // result unused -> REDG.E.ADD.F32
            if (b.bin + 1 < nb) atomicAdd(cell + planes_per_bin * plane, wr);
        }
    }
}

__global__ void __launch_bounds__(kScatterThreads)
voxel_scatter_atomic_kernel(const double *__restrict__ ev, const int64_t *__restrict__ off,
                            int B, int nb, int H, int W, int flavour, float *__restrict__ out, int b_first, int b_end) {
    scatter_body(ev, off, B, nb, H, W, flavour, out, b_first, b_end, (int)blockIdx.x, (int)gridDim.x);
}

// ---- scatter + statistics in one pass (std normalisation) ------------------------------------------------
// The statistics pass of the L2 path reads every grid of the chunk once more (~100 us of the 370 at 64 x 480x640) only to
// form sum, sum of squares and the count of non-zero cells of the hot-pixel-filtered grid.  With RETURNING atomics they
// telescope out of the scatter: an add takes a cell from `old` to `old + w` (the SM repeats the L2's round-to-nearest add;
// no subnormals occur: |w| >= 2^-53), and f(new) - f(old), f(new)^2 - f(old)^2, [f(new) != 0] - [f(old) != 0] summed over all
// adds of a cell give f(final), f(final)^2 and [f(final) != 0] (f = the hot-pixel filter; every difference and square is
// exact in fp64).  Returning atomics retire at 87 G/s against 105 G/s for RED (profiles/r02/atomics_l2_probe.txt): the
// scatter gets ~20 % slower and the statistics pass disappears.  Per thread fp64 accumulators, reduced per warp, three
// atomics per warp and window into the window's first Partial (zeroed by the host); the normalise kernel is unchanged.
// MEASURED (B200, same box, us per launch incl. normalisation): 64 x 480x640 419 against 367 with the separate statistics pass,
// 8 x 480x640 50 against 44, 1 x 624x970 (1 M events) 56 against 30, 64 x 180x240 54.5 against 52.5 -- the returned values
// make every add a round trip the thread waits for (8 in flight per thread at 2 CTAs per SM), far below the 87 G/s the
// probe reached with nothing else to do.  Kept as an experiment (CF_VOXEL_FLAGS bit5), parity-tested; not the default.
struct TelAcc {
    double s, q;
    long long n;
    int b;
};
__device__ __forceinline__ void tel_flush(TelAcc &a, Partial *__restrict__ tel, int tel_stride) {
    // lanes of a warp nearly always hold the same window: one reduced update then, else every lane for itself
    const unsigned full = 0xffffffffu;
    const int b0 = __shfl_sync(full, a.b, 0);
    if (__all_sync(full, a.b == b0)) {
        const double s = warp_sum(a.s), q = warp_sum(a.q);
        const long long n = warp_sum(a.n);
        if ((threadIdx.x & 31) == 0 && b0 >= 0 && (s != 0.0 || q != 0.0 || n != 0)) {
            Partial *p = tel + (size_t)b0 * tel_stride;
            atomicAdd(&p->sum, s);
            atomicAdd(&p->sumsq, q);
            atomicAdd(reinterpret_cast<unsigned long long *>(&p->nnz), (unsigned long long)n);
        }
    } else if (a.b >= 0 && (a.s != 0.0 || a.q != 0.0 || a.n != 0)) {
        Partial *p = tel + (size_t)a.b * tel_stride;
        atomicAdd(&p->sum, a.s);
        atomicAdd(&p->sumsq, a.q);
        atomicAdd(reinterpret_cast<unsigned long long *>(&p->nnz), (unsigned long long)a.n);
    }
    a.s = 0.0; a.q = 0.0; a.n = 0;
}

__global__ void __launch_bounds__(kScatterThreads)
voxel_scatter_stats_kernel(const double *__restrict__ ev, const int64_t *__restrict__ off, int B, int nb, int H, int W,
                           int flavour, float *__restrict__ out, int b_first, int b_end, float hot_thr,
                           Partial *__restrict__ tel, int tel_stride) {
    const int blk = (int)blockIdx.x, nblk = (int)gridDim.x;
    const int64_t ev_first = __ldg(off + b_first), ev_end = __ldg(off + b_end);
    const int64_t plane = (int64_t)H * W;
    const int planes_per_bin = flavour == CF_FLAVOUR_POL ? 2 : 1;
    constexpr int64_t kTileEvents = kScatterThreads * kScatterUnroll;
    Window w;
    w.b = -1;
    TelAcc acc{0.0, 0.0, 0, -1};
    for (int64_t tile = ev_first + (int64_t)blk * kTileEvents; tile < ev_end; tile += (int64_t)nblk * kTileEvents) {
        Event e[kScatterUnroll];
#pragma unroll
        for (int k = 0; k < kScatterUnroll; ++k) {  // all loads in flight first
            const int64_t i = tile + k * kScatterThreads + threadIdx.x;
            if (i < ev_end) e[k] = load_event(ev, i);
        }
        // all adds of the tile in flight before the first returned value is used
        float *cell[kScatterUnroll];
        float wl[kScatterUnroll], wr[kScatterUnroll], ol[kScatterUnroll], orr[kScatterUnroll];
        int wb[kScatterUnroll];
        bool two[kScatterUnroll];
#pragma unroll
        for (int k = 0; k < kScatterUnroll; ++k) {
            const int64_t i = tile + k * kScatterThreads + threadIdx.x;
            cell[k] = nullptr;
            two[k] = false;
            wb[k] = -1;
            if (i >= ev_end) continue;
            locate_window(w, i, off, ev, B);
            const Binned b = bin_event(e[k], w, nb, H, W, flavour);
            if (!b.ok) continue;
            if (flavour == CF_FLAVOUR_TORCH) {
                weights_f32(b, wl[k], wr[k]);
            } else {
                double dl, dr;
                weights_wide(b, flavour, dl, dr);
                wl[k] = (float)dl;
                wr[k] = (float)dr;
            }
            cell[k] = out + (((int64_t)w.b * nb + b.bin) * planes_per_bin + b.chan) * plane + (int64_t)b.y * W + b.x;
            two[k] = b.bin + 1 < nb;
            wb[k] = w.b;
            ol[k] = atomicAdd(cell[k], wl[k]);
            if (two[k]) orr[k] = atomicAdd(cell[k] + planes_per_bin * plane, wr[k]);
        }
#pragma unroll
        for (int k = 0; k < kScatterUnroll; ++k) {
            const bool live = cell[k] != nullptr;
            // (warp-uniform test first: the flush is a warp collective)
            if (__any_sync(0xffffffffu, live && wb[k] != acc.b && acc.b >= 0)) tel_flush(acc, tel, tel_stride);
            if (!live) continue;
            acc.b = wb[k];
            auto take = [&](float old, float add) {
                const float fo = hot_filter(old, hot_thr), fn = hot_filter(__fadd_rn(old, add), hot_thr);
                acc.s += (double)fn - (double)fo;
                acc.q += (double)fn * (double)fn - (double)fo * (double)fo;
                acc.n += (long long)(fn != 0.f) - (long long)(fo != 0.f);
            };
            take(ol[k], wl[k]);
            if (two[k]) take(orr[k], wr[k]);
        }
    }
    tel_flush(acc, tel, tel_stride);
}

// ----------------------------------------- atomic mode, cluster (smem) path ---
// fp32 add into the shared memory of CTA `rank` of this cluster (DSMEM), no return value
__device__ __forceinline__ void red_add_cluster(float *local_ptr, unsigned rank, float v) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(local_ptr);
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    asm volatile("red.relaxed.cluster.shared::cluster.add.f32 [%0], %1;" ::"r"(r), "f"(v) : "memory");
}

constexpr int kClusterThreads = 256;

__global__ void __launch_bounds__(kClusterThreads)
voxel_cluster_kernel(const double *__restrict__ ev, const int64_t *__restrict__ off, int B, int nb, int H, int W,
                     int flavour, int preprocess, float hot_thr, float *__restrict__ out, int slice) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank(), CS = cluster.num_blocks();
    extern __shared__ __align__(16) float sm[];          // [slice] grid slice, then one Partial
    Partial *part = reinterpret_cast<Partial *>(sm + slice);
    __shared__ Partial warp_part[kClusterThreads / 32];
    __shared__ double s_a, s_inv;
    __shared__ int s_identity;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int planes = flavour == CF_FLAVOUR_POL ? 2 : 1;
    const int64_t plane = (int64_t)H * W;
    const int cells = (int)((int64_t)nb * planes * plane);
    const int my_first = (int)rank * slice;
    const int valid = max(0, min(slice, cells - my_first));  // cells this CTA owns
    const int n_clusters = gridDim.x / CS;

    for (int b = blockIdx.x / CS; b < B; b += n_clusters) {
        // ---- phase 0: clear the slice
        for (int i = tid; i < slice / 4; i += kClusterThreads) reinterpret_cast<float4 *>(sm)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        cluster.sync();

        // ---- phase 1: this CTA's share of the window's events -> DSMEM scatter
        const int64_t begin = __ldg(off + b), end = __ldg(off + b + 1);
        if (end > begin) {
            Window w;
            w.b = b; w.begin = begin; w.end = end;
            w.t0 = __ldg(ev + 4 * begin);
            w.span = __dsub_rn(__ldg(ev + 4 * (end - 1)), w.t0);
            if (w.span == 0.0) w.span = 1.0;
            const int64_t per = (end - begin + CS - 1) / CS;
            const int64_t s = begin + (int64_t)rank * per, e = min(end, s + per);
            constexpr int U = 4;
            for (int64_t i0 = s + tid; i0 < e; i0 += (int64_t)U * kClusterThreads) {
                Event evs[U];
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const int64_t i = i0 + (int64_t)k * kClusterThreads;
                    if (i < e) evs[k] = load_event(ev, i);
                }
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const int64_t i = i0 + (int64_t)k * kClusterThreads;
                    if (i >= e) break;
                    const Binned bb = bin_event(evs[k], w, nb, H, W, flavour);
                    if (!bb.ok) continue;
                    float wl, wr;
                    if (flavour == CF_FLAVOUR_TORCH) {
                        weights_f32(bb, wl, wr);
                    } else {
                        double dl, dr;
                        weights_wide(bb, flavour, dl, dr);
                        wl = (float)dl;
                        wr = (float)dr;
                    }
                    const int cell = (int)(((int64_t)bb.bin * planes + bb.chan) * plane + (int64_t)bb.y * W + bb.x);
                    const unsigned own = (unsigned)cell / (unsigned)slice;
                    red_add_cluster(sm + (cell - (int)own * slice), own, wl);
                    if (bb.bin + 1 < nb) {
                        const int cell2 = cell + (int)(planes * plane);
                        const unsigned own2 = (unsigned)cell2 / (unsigned)slice;
                        red_add_cluster(sm + (cell2 - (int)own2 * slice), own2, wr);
                    }
                }
            }
        }
        cluster.sync();

        // ---- phase 2: statistics of the slice, exchanged through DSMEM (fixed order => every
        //      CTA of the cluster derives bit-identical mean / scale)
        double a = 0.0, inv = 1.0;
        bool identity = true;
        if (preprocess != CF_PRE_NONE) {
            double sum = 0.0, sumsq = 0.0;
            long long nnz = 0;
            float mn = INFINITY, mx = -INFINITY;
            for (int i = tid; i < valid; i += kClusterThreads) {
                const float v = hot_filter(sm[i], hot_thr);
                sum += (double)v;
                sumsq += (double)v * (double)v;
                nnz += (v != 0.f);
                mn = fminf(mn, v);
                mx = fmaxf(mx, v);
            }
            sum = warp_sum(sum); sumsq = warp_sum(sumsq); nnz = warp_sum(nnz);
            mn = warp_min(mn); mx = warp_max(mx);
            if (lane == 0) warp_part[warp] = Partial{sum, sumsq, nnz, mn, mx};
            __syncthreads();
            if (tid == 0) {
                Partial t = warp_part[0];
                for (int k = 1; k < kClusterThreads / 32; ++k) {
                    t.sum += warp_part[k].sum; t.sumsq += warp_part[k].sumsq; t.nnz += warp_part[k].nnz;
                    t.mn = fminf(t.mn, warp_part[k].mn); t.mx = fmaxf(t.mx, warp_part[k].mx);
                }
                *part = t;
            }
            cluster.sync();
            if (warp == 0) {
                double ts = 0.0, tq = 0.0;
                long long tn = 0;
                float tmn = INFINITY, tmx = -INFINITY;
                if ((unsigned)lane < CS) {
                    const Partial *rp = cluster.map_shared_rank(part, lane);
                    ts = rp->sum; tq = rp->sumsq; tn = rp->nnz; tmn = rp->mn; tmx = rp->mx;
                }
                ts = warp_sum(ts); tq = warp_sum(tq); tn = warp_sum(tn);
                tmn = warp_min(tmn); tmx = warp_max(tmx);
                if (lane == 0) {
                    if (preprocess == CF_PRE_STD) {
                        s_identity = tn == 0;
                        const double mean = tn ? ts / (double)tn : 0.0;
                        const double var = tn ? tq / (double)tn - mean * mean : 0.0;
                        s_a = mean;
                        s_inv = 1.0 / (sqrt(fmax(var, 0.0)) + 1e-8);
                    } else {
                        s_identity = 0;
                        s_a = (double)tmn;
                        s_inv = 1.0 / ((double)tmx - (double)tmn + 1e-8);
                    }
                }
            }
            __syncthreads();
            a = s_a; inv = s_inv; identity = s_identity != 0;
        }

        // ---- phase 3: normalise the slice and write it once
        auto norm = [&](float raw) -> float {
            if (preprocess == CF_PRE_NONE) return raw;
            const float v = hot_filter(raw, hot_thr);
            if (identity) return v;
            if (preprocess == CF_PRE_STD) return (v != 0.f) ? (float)(((double)v - a) * inv) : 0.f;
            return (float)(((double)v - a) * inv);
        };
        float *o = out + (int64_t)b * cells + my_first;
        if ((cells & 3) == 0) {
            for (int i = tid; i < valid / 4; i += kClusterThreads) {
                const float4 q = reinterpret_cast<const float4 *>(sm)[i];
                st_cs4(reinterpret_cast<float4 *>(o) + i, make_float4(norm(q.x), norm(q.y), norm(q.z), norm(q.w)));
            }
        } else {
            for (int i = tid; i < valid; i += kClusterThreads) st_cs(o + i, norm(sm[i]));
        }
        // the next window's phase-0 cluster.sync() orders these reads before any remote add
        __syncthreads();
    }
    cluster.sync();  // no CTA may exit while a peer can still address its shared memory
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
    if (i >= __ldg(off + B)) {   // rows past the last window (a device-side filter kept fewer rows than the buffer holds)
        keys[i] = invalid_key;
        vals[i] = (uint32_t)i;
        return;
    }
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
// event order; for time-sorted events bin indices are non-decreasing along it (runs that are not take
// the order-agnostic walk at the top of the kernel).  For each bin: first
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

    // The fast walk below needs the run's bin indices to be non-decreasing (time-sorted events).  The reference's
    // scatter-adds do not depend on that, so a run that breaks it (jittered / unsorted stamps, concatenated packs)
    // takes the order-agnostic walk: per bin, the whole run in event order.
    {
        int64_t e = i;
        int last_bin = -1;
        bool monotone = true;
        while (e < n && keys[e] == k) {
            const Binned bb = bin_event(load_event(ev, vals[e]), w, nb, H, W, flavour);
            monotone = monotone && bb.bin >= last_bin;
            last_bin = bb.bin;
            ++e;
        }
        if (!monotone) {
            for (int bin = 0; bin < nb; ++bin) {
                float acc = 0.f;
                bool touched = false;
                if (flavour == CF_FLAVOUR_MVSEC) {  // one index_put_ per bin: left and right weights interleave in event order
                    for (int64_t j = i; j < e; ++j) {
                        const Binned bb = bin_event(load_event(ev, vals[j]), w, nb, H, W, flavour);
                        if (bb.bin != bin && bb.bin != bin - 1) continue;
                        double wl, wr;
                        weights_mvsec(bb, wl, wr);
                        acc = __fadd_rn(acc, (float)(bb.bin == bin ? wl : wr));
                        touched = true;
                    }
                } else {
                    for (int pass = 0; pass < 2; ++pass) {  // all left weights (ti == bin), then all right ones (ti == bin-1)
                        for (int64_t j = i; j < e; ++j) {
                            const Binned bb = bin_event(load_event(ev, vals[j]), w, nb, H, W, flavour);
                            if (bb.bin != bin - pass) continue;
                            if (flavour == CF_FLAVOUR_TORCH) {
                                float wl, wr;
                                weights_f32(bb, wl, wr);
                                acc = __fadd_rn(acc, pass ? wr : wl);
                            } else {
                                double wl, wr;
                                weights_f64(bb, wl, wr);
                                acc = (float)__dadd_rn((double)acc, pass ? wr : wl);
                            }
                            touched = true;
                        }
                    }
                }
                if (touched) cell0[(int64_t)bin * planes * plane] = acc;
            }
            return;
        }
    }

    int64_t pos = i, prev_s = i, prev_e = i;
    for (int bin = 0; bin < nb; ++bin) {
        float acc = 0.f;
        bool touched = false;
        int64_t e = pos;
        if (flavour == CF_FLAVOUR_MVSEC) {
            // one index_put_ per bin over the time-sorted events (MVSEC_utils.py:283-292): the events of bin-1 (their
            // right weights) precede the events of this bin (their left weights); fp32 adds of fp32-cast weights
            for (int64_t j = prev_s; j < prev_e; ++j) {
                const Binned bb = bin_event(load_event(ev, vals[j]), w, nb, H, W, flavour);
                double wl, wr;
                weights_mvsec(bb, wl, wr);
                acc = __fadd_rn(acc, (float)wr);
                touched = true;
            }
            while (e < n && keys[e] == k) {
                const Binned bb = bin_event(load_event(ev, vals[e]), w, nb, H, W, flavour);
                if (bb.bin != bin) break;
                double wl, wr;
                weights_mvsec(bb, wl, wr);
                acc = __fadd_rn(acc, (float)wl);
                touched = true;
                ++e;
            }
            if (touched) cell0[(int64_t)bin * planes * plane] = acc;
            prev_s = pos;
            prev_e = e;
            pos = e;
            if (prev_s == prev_e && (pos >= n || keys[pos] != k)) break;
            continue;
        }
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
constexpr int kStatThreads = 256;

// Both kernels: one CTA = one chunk of a window; a thread streams its float4s with 4 loads in flight.
// (Round-1a sized them for 296 CTAs with one load in flight per thread: 43 of 70 us at 8 x 480x640.)
__device__ __forceinline__ void stats_body(const float *__restrict__ grid, int64_t cells, int64_t chunk_len, float hot_thr,
                                           Partial *__restrict__ partials, int chunks, int b, int c) {
    const float *g = grid + (int64_t)b * cells;
    const int64_t s = (int64_t)c * chunk_len, e = min(cells, s + chunk_len);
    // short fp32 partials per thread (a few dozen terms), promoted to fp64 across threads and chunks
    float fs = 0.f, fq = 0.f;
    int fn = 0;
    float mn = INFINITY, mx = -INFINITY;
    auto take = [&](float raw) {
        const float v = hot_filter(raw, hot_thr);
        fs += v;
        fq = fmaf(v, v, fq);
        fn += (v != 0.f);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    };
    if (((cells | chunk_len) & 3) == 0) {  // 128-bit loads (chunk starts stay 16-byte aligned)
        const float4 *g4 = reinterpret_cast<const float4 *>(g);
        const int64_t e4 = e >> 2;
        for (int64_t i = (s >> 2) + threadIdx.x; i < e4; i += 4 * kStatThreads) {
            float4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t j = i + (int64_t)u * kStatThreads;
                q[u] = j < e4 ? __ldg(g4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i + (int64_t)u * kStatThreads < e4) { take(q[u].x); take(q[u].y); take(q[u].z); take(q[u].w); }
            }
        }
    } else {
        for (int64_t i = s + threadIdx.x; i < e; i += kStatThreads) take(__ldg(g + i));
    }
    // fixed-shape tree: deterministic for a given launch geometry
    double sum = warp_sum((double)fs), sumsq = warp_sum((double)fq);
    long long nnz = warp_sum((long long)fn);
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
voxel_stats_kernel(const float *__restrict__ grid, int64_t cells, int64_t chunk_len, float hot_thr,
                   Partial *__restrict__ partials, int chunks) {
    stats_body(grid, cells, chunk_len, hot_thr, partials, chunks, (int)blockIdx.y, (int)blockIdx.x);
}

__device__ __forceinline__ void normalise_body(const float *in, float *out /* may alias in */, int64_t cells, int64_t chunk_len,
                                               float hot_thr, int mode, const Partial *__restrict__ partials, int chunks,
                                               int b, int c) {
    const float *g = in + (int64_t)b * cells;
    float *o = out + (int64_t)b * cells;
    const int64_t s = (int64_t)c * chunk_len, e = min(cells, s + chunk_len);
    const bool vec = ((cells | chunk_len) & 3) == 0;
    // this thread's first loads go out before the window's partials are reduced
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    float4 *o4 = reinterpret_cast<float4 *>(o);
    const int64_t e4 = e >> 2, i0 = (s >> 2) + threadIdx.x;
    float4 q[4];
    if (vec) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t j = i0 + (int64_t)u * kStatThreads;
            q[u] = j < e4 ? g4[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __shared__ float s_a, s_b;   // out = (v - s_a) * s_b
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
                s_a = (float)mean;
                s_b = (float)(1.0 / (sqrt(fmax(var, 0.0)) + 1e-8));
            } else {
                s_identity = 0;
                s_a = mn;
                s_b = (float)(1.0 / ((double)mx - (double)mn + 1e-8));
            }
        }
    }
    __syncthreads();
    const float a = s_a, inv = s_b;
    const bool identity = s_identity != 0;
    // mean and 1/(std+eps) are formed in fp64 once per window; the per-cell map is fp32 (1 ulp of the
    // fp64 formula rounded to fp32; a true fp64 divide per cell made this kernel 3x slower than the scatter)
    auto norm = [&](float raw) -> float {
        const float v = hot_filter(raw, hot_thr);
        if (identity) return v;
        const float r = (v - a) * inv;
        return (mode == CF_PRE_STD && v == 0.f) ? 0.f : r;
    };
    // in place, std mode: zero cells stay zero (event_process.py:207-210) -- an all-zero float4 is not rewritten
    const bool skip_zero = (in == out) && mode == CF_PRE_STD;
    if (vec) {
        for (int64_t i = i0; i < e4; i += 4 * kStatThreads) {
            if (i != i0) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t j = i + (int64_t)u * kStatThreads;
                    q[u] = j < e4 ? g4[j] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t j = i + (int64_t)u * kStatThreads;
                if (j < e4) {
                    if (skip_zero && q[u].x == 0.f && q[u].y == 0.f && q[u].z == 0.f && q[u].w == 0.f) continue;
                    o4[j] = make_float4(norm(q[u].x), norm(q[u].y), norm(q[u].z), norm(q[u].w));
                }
            }
        }
    } else {
        for (int64_t i = s + threadIdx.x; i < e; i += kStatThreads) o[i] = norm(g[i]);
    }
}

__global__ void __launch_bounds__(kStatThreads)
voxel_normalise_kernel(const float *in, float *out /* may alias in */, int64_t cells, int64_t chunk_len,
                       float hot_thr, int mode, const Partial *__restrict__ partials, int chunks) {
    normalise_body(in, out, cells, chunk_len, hot_thr, mode, partials, chunks, (int)blockIdx.y, (int)blockIdx.x);
}

// ---------------------------------------------------------- pipelined launches ---
// The four stages of the L2-atomic path (zero, scatter, statistics, normalise) of FOUR consecutive window chunks in ONE
// launch: CTAs [0, n_scatter) scatter chunk k, the next n_stats CTAs reduce chunk k-1, then chunk k-2 is normalised and
// chunk k+1 zero-filled.  The stages bind on different units -- the scatter on the L2's atomic ALUs (~105 G RED/s,
// profiles/r02/atomics_l2_probe.txt), the other three on L2 bandwidth -- and run back to back each of them left the
// machine half idle plus a launch ramp and tail per kernel (4 launches per chunk).  B + 3 launches of this kernel
// replace 4 B of the single-stage kernels; all chunks in flight stay L2-resident (the host sizes them).
struct PipeStage {
    int b0, nbat, ctas;   // windows [b0, b0 + nbat) of the batch, CTAs of this stage (0: idle)
};
__global__ void __launch_bounds__(kScatterThreads)
voxel_pipeline_kernel(const double *__restrict__ ev, const int64_t *__restrict__ off, int B, int nb, int H, int W, int flavour,
                      float *__restrict__ out, int64_t cells, int64_t chunk_len, int chunks, float hot_thr, int mode,
                      Partial *__restrict__ partials, PipeStage scat, PipeStage stat, PipeStage norm, PipeStage zero) {
    static_assert(kScatterThreads == kStatThreads, "one block size for all stages");
    int blk = (int)blockIdx.x;
    if (blk < scat.ctas) {
        scatter_body(ev, off, B, nb, H, W, flavour, out, scat.b0, scat.b0 + scat.nbat, blk, scat.ctas);
        return;
    }
    blk -= scat.ctas;
    if (blk < stat.ctas) {
        stats_body(out, cells, chunk_len, hot_thr, partials, chunks, stat.b0 + blk / chunks, blk % chunks);
        return;
    }
    blk -= stat.ctas;
    if (blk < norm.ctas) {
        normalise_body(out, out, cells, chunk_len, hot_thr, mode, partials, chunks, norm.b0 + blk / chunks, blk % chunks);
        return;
    }
    blk -= norm.ctas;
    {
        float4 *z = reinterpret_cast<float4 *>(out + (int64_t)zero.b0 * cells);   // cells % 4 == 0 (checked by the host)
        const int64_t n4 = (int64_t)zero.nbat * cells / 4;
        for (int64_t i = (int64_t)blk * kScatterThreads + threadIdx.x; i < n4; i += (int64_t)zero.ctas * kScatterThreads)
            z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

static void stat_geometry(int B, int64_t cells, int &chunks, int64_t &chunk_len) {
    // ~8 CTAs per SM over the batch, chunks of >= 2048 cells (2 float4 per thread), <= kMaxChunks per window
    int64_t want = ceil_div(8 * 148, B > 0 ? B : 1);
    int64_t cap = ceil_div(cells, 2048);
    chunks = (int)(want < 1 ? 1 : want);
    if (chunks > cap) chunks = (int)(cap < 1 ? 1 : cap);
    if (chunks > kMaxChunks) chunks = kMaxChunks;
    chunk_len = ceil_div(cells, chunks);
    chunk_len = (chunk_len + 3) & ~(int64_t)3;  // keeps every chunk start 16-byte aligned
}

static int run_preprocess(const float *in, float *out, int B, int64_t cells, int preprocess, float hot_thr,
                          void *ws, size_t ws_bytes, cudaStream_t stream, bool stats_done = false) {
    CF_REQUIRE(B <= 65535, CF_ERR_INVALID_ARG, "preprocess: B > 65535");
    CF_REQUIRE(ws && ws_bytes >= (size_t)B * kMaxChunks * sizeof(Partial), CF_ERR_WORKSPACE,
               "preprocess: workspace too small (%zu < %zu)", ws_bytes, (size_t)B * kMaxChunks * sizeof(Partial));
    CF_REQUIRE(aligned16(ws), CF_ERR_ALIGN, "preprocess: workspace not 16-byte aligned");
    int chunks;
    int64_t chunk_len;
    stat_geometry(B, cells, chunks, chunk_len);
    Partial *partials = reinterpret_cast<Partial *>(ws);
    dim3 grid(chunks, B);
    if (!stats_done) {   // (else the scatter left each window's totals in its first Partial, the others zero)
        voxel_stats_kernel<<<grid, kStatThreads, 0, stream>>>(in, cells, chunk_len, hot_thr, partials, chunks);
        CF_LAUNCH_CHECK("voxel_stats_kernel");
    }
    voxel_normalise_kernel<<<grid, kStatThreads, 0, stream>>>(in, out, cells, chunk_len, hot_thr, preprocess, partials, chunks);
    CF_LAUNCH_CHECK("voxel_normalise_kernel");
    return CF_OK;
}

// statistics + normalisation for the packed-event path (voxel_packed.cu)
int run_preprocess_shared(const float *in, float *out, int B, int64_t cells, int preprocess, float hot_thr, void *ws,
                          size_t ws_bytes, cudaStream_t stream) {
    return run_preprocess(in, out, B, cells, preprocess, hot_thr, ws, ws_bytes, stream);
}

// implemented in voxel_tiled.cu (CF_VOXEL_ATOMIC_TILED)
size_t voxel_tiled_workspace_bytes(int64_t total, int B, int nb, int H, int W, int flavour);
int launch_voxel_tiled(const double *events, const int64_t *offsets, int64_t total, int B, int nb, int H, int W,
                       int flavour, int preprocess, float hot_thr, float *out, void *ws, size_t ws_bytes,
                       cudaStream_t stream);

// ATOMIC mode has two data paths with identical numerics (sum order unspecified in both):
//   L2 atomics (this file, the default): zero -> RED.ADD into the L2-resident grid -> statistics ->
//     normalise in place;
//   tiled (voxel_tiled.cu, CF_VOXEL_ATOMIC_TILED): partition into 12-byte records -> accumulate tiles in
//     shared memory -> statistics, normalisation and the only write of the grid out of shared memory.
// Measured on the B200 (scripts/scale_bench.py, us per launch, fused std normalisation, round 1):
//                 8x180x240  64x180x240  1x260x346  64x260x346  8x480x640  1x624x970
//   L2 atomics       16.4       50.8       10.3       132.0       43.8       30.0
//   tiled            19.4      107.5       15.4       160.1       57.1       49.6
// The tiled path issues no global atomic, but its two kernels are each a chain of dependent steps
// (window lookup -> event loads -> fp64 decode -> counting sort -> write; run bounds -> records ->
// shared-memory atomics -> statistics -> per-window barrier -> normalise -> write: timelines from
// scripts/voxel_trace.py in profiles/), two waves are needed as soon as the batch's grids exceed the
// chip's 33 MB of shared memory, and the ~100 G/s of L2 atomics it avoids were not the bound once the
// statistics/normalise kernels were given enough loads in flight.  So the L2 path is the default and the
// tiled path stays selectable (and tested) for the next round's work on it.
static int voxel_flags();
static bool use_tiled(int64_t, int, int64_t) {
    return (voxel_flags() & 8) != 0;              // experiments: CF_VOXEL_FLAGS=8 forces the tiled path
}

// CF_VOXEL_FLAGS (debug / experiments): bit5 = statistics telescoped out of a returning-atomics scatter, bit4 = pipelined launches (four stages of four chunks per launch), bit3 = force the tiled path,
// bit2 = force the L2-atomic path, bit1 = use the cluster / DSMEM-atomics path, bit0 = do not chunk
static int voxel_flags() {
    static int flags = -1;
    if (flags < 0) {
        const char *e = getenv("CF_VOXEL_FLAGS");
        flags = e ? atoi(e) : 0;
    }
    return flags;
}

// Returns CF_OK / an error, or 1 when the window grid does not fit a 16-CTA cluster's shared memory.
static int launch_cluster_path(const double *events, const int64_t *offsets, int B, int nb, int H, int W, int flavour,
                               int preprocess, float hot_thr, float *out, int64_t cells, cudaStream_t stream) {
    constexpr size_t kMaxSlice = 200 * 1024;  // bytes of grid per CTA
    if (cells * sizeof(float) > 16 * kMaxSlice || cells >= (1ll << 30)) return 1;
    // smallest power-of-two cluster that holds the grid; grow it while the launch would leave SMs idle
    int cs = 1;
    while ((size_t)ceil_div(cells, cs) * sizeof(float) > kMaxSlice) cs *= 2;
    const int sms = sm_count();
    while (cs < 16 && (int64_t)B * cs * 2 <= sms && ceil_div(cells, cs * 2) >= 4096) cs *= 2;
    if (const char *force = getenv("CF_VOXEL_CS")) {  // experiments: force the cluster size (must still fit)
        const int f = atoi(force);
        if (f >= cs && f <= 16 && (f & (f - 1)) == 0) cs = f; else if (f > 0 && f < cs) return 1;
    }
    const int slice = (int)((ceil_div(cells, cs) + 3) & ~(int64_t)3);
    const size_t smem = (size_t)slice * sizeof(float) + sizeof(Partial);
    int dev = 0;
    CF_CUDA(cudaGetDevice(&dev));
    static bool configured[64] = {};
    if (!configured[dev & 63]) {
        CF_CUDA(cudaFuncSetAttribute(voxel_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kMaxSlice + 1024)));
        CF_CUDA(cudaFuncSetAttribute(voxel_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        configured[dev & 63] = true;
    }
    const int per_sm = (int)((227 * 1024) / (smem + 1024));
    int64_t max_clusters = (int64_t)(sms / cs) * (per_sm < 1 ? 1 : per_sm);
    if (cs == 16) max_clusters = 8 * (per_sm < 1 ? 1 : per_sm);  // one 16-CTA cluster per GPC
    if (max_clusters < 1) max_clusters = 1;
    const int n_clusters = (int)(B < max_clusters ? B : max_clusters);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_clusters * cs));
    cfg.blockDim = dim3(kClusterThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, voxel_cluster_kernel, events, offsets, B, nb, H, W, flavour, preprocess, hot_thr,
                                       out, slice);
    count_launch("voxel_cluster_kernel");
    if (e != cudaSuccess) {
        set_error("launch of voxel_cluster_kernel (cluster %d, %zu B smem) failed: %s", cs, smem, cudaGetErrorString(e));
        return CF_ERR_CUDA;
    }
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

extern "C" size_t cf_voxel_workspace_bytes(int64_t total_events, int B, int nb, int H, int W, int mode, int flavour, int) {
    size_t base = cf::align_up(cf_preprocess_workspace_bytes(B, 0), 256);
    if (mode == CF_VOXEL_DETERMINISTIC) return cf::det_layout(total_events, base).end;
    if (mode != CF_VOXEL_ATOMIC_L2 && nb > 0 && H > 0 && W > 0)
        base += cf::voxel_tiled_workspace_bytes(total_events, B, nb, H, W, flavour);
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
    CF_REQUIRE(mode >= CF_VOXEL_ATOMIC && mode <= CF_VOXEL_ATOMIC_TILED, CF_ERR_INVALID_ARG, "cf_voxel_bin: bad mode %d", mode);
    const bool force_l2 = mode == CF_VOXEL_ATOMIC_L2, force_tiled = mode == CF_VOXEL_ATOMIC_TILED;
    if (force_l2 || force_tiled) mode = CF_VOXEL_ATOMIC;
    CF_REQUIRE(flavour >= CF_FLAVOUR_TORCH && flavour <= CF_FLAVOUR_MVSEC, CF_ERR_INVALID_ARG, "cf_voxel_bin: bad flavour %d", flavour);
    CF_REQUIRE(!(force_tiled && flavour == CF_FLAVOUR_MVSEC), CF_ERR_UNSUPPORTED,
               "cf_voxel_bin: CF_VOXEL_ATOMIC_TILED does not implement CF_FLAVOUR_MVSEC");
    CF_REQUIRE(preprocess >= CF_PRE_NONE && preprocess <= CF_PRE_MAXMIN, CF_ERR_INVALID_ARG, "cf_voxel_bin: bad preprocess %d", preprocess);
    CF_REQUIRE(total == 0 || aligned16(events), CF_ERR_ALIGN, "cf_voxel_bin: events not 16-byte aligned");
    CF_REQUIRE(total < (1ll << 32), CF_ERR_INVALID_ARG, "cf_voxel_bin: more than 2^32 events in one call");
    if (B == 0) return CF_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int planes = flavour == CF_FLAVOUR_POL ? 2 : 1;
    const int64_t cells = (int64_t)nb * planes * H * W;
    if (mode == CF_VOXEL_ATOMIC && (voxel_flags() & 2)) {  // experimental, see the file header
        int rc = launch_cluster_path(events, offsets, B, nb, H, W, flavour, preprocess, hot_thr, out, cells, stream);
        if (rc != 1) return rc;  // 1 = grid too large for the cluster path -> global path below
    }
    if (mode == CF_VOXEL_ATOMIC && !force_l2 && !(voxel_flags() & 4) && flavour != CF_FLAVOUR_MVSEC &&
        (force_tiled || use_tiled(total, B, (int64_t)nb * planes * H * W))) {
        const size_t base = align_up(cf_preprocess_workspace_bytes(B, 0), 256);
        const size_t need = voxel_tiled_workspace_bytes(total, B, nb, H, W, flavour);
        if (need > 0 && total > 0) {
            CF_REQUIRE(ws && ws_bytes >= base + need, CF_ERR_WORKSPACE, "cf_voxel_bin: workspace too small (%zu < %zu)",
                       ws_bytes, base + need);
            const int rc = launch_voxel_tiled(events, offsets, total, B, nb, H, W, flavour, preprocess, hot_thr, out,
                                              reinterpret_cast<char *>(ws) + base, ws_bytes - base, stream);
            if (rc != 1) return rc;  // 1 = geometry not covered -> L2-atomic path below
        }
        CF_REQUIRE(!force_tiled || total == 0, CF_ERR_UNSUPPORTED,
                   "cf_voxel_bin: CF_VOXEL_ATOMIC_TILED does not cover B=%d, %dx%d, nb=%d", B, H, W, nb);
    }
    if (mode == CF_VOXEL_ATOMIC) {
        // Chunks of windows that stay L2-resident from the zero-fill to the normalised write-back.
        // The host does not know the per-window event counts (offsets live on the device), so the
        // events of a chunk are budgeted with the batch average; kernels take device offsets.
        const size_t per_window = (size_t)cells * sizeof(float) + (size_t)(total / B + 1) * 32;
        static const size_t budget_mb = [] { const char *e = getenv("CF_VOXEL_CHUNK_MB"); return e ? (size_t)atoi(e) : (size_t)96; }();
        int chunk = (voxel_flags() & 1) ? B : (int)((budget_mb << 20) / per_window);
        if (chunk < 1) chunk = 1;
        if (chunk > B) chunk = B;
        if (preprocess != CF_PRE_NONE) {
            CF_REQUIRE(ws && ws_bytes >= (size_t)B * kMaxChunks * sizeof(Partial), CF_ERR_WORKSPACE,
                       "cf_voxel_bin: workspace too small (%zu < %zu)", ws_bytes, (size_t)B * kMaxChunks * sizeof(Partial));
            CF_REQUIRE(aligned16(ws), CF_ERR_ALIGN, "cf_voxel_bin: workspace not 16-byte aligned");
        }
        if ((voxel_flags() & 16) && (cells & 3) == 0 && total > 0 && aligned16(out)) {
            // ---- pipelined (experiment, CF_VOXEL_FLAGS bit4; measured SLOWER than one stage per launch on the B200: 64 x 480x640
            //      573 against 365 us, 8 x 480x640 78 against 43 us -- the stages do not overlap usefully inside one launch):
            //      stage s of chunk k runs in launch k + s (voxel_pipeline_kernel).  Four chunks are in
            //      flight (zeroed | being scattered into | being reduced | being normalised): size them so that all
            //      four stay L2-resident
            int pc = (voxel_flags() & 1) ? B : (int)((budget_mb << 20) / (4 * per_window));
            if (pc < 1) pc = 1;
            if (pc > B) pc = B;
            const int nch = (int)ceil_div(B, pc);
            int chunks = 1;
            int64_t chunk_len = cells;
            if (preprocess != CF_PRE_NONE) stat_geometry(pc, cells, chunks, chunk_len);
            Partial *partials = reinterpret_cast<Partial *>(ws);
            const int64_t cap = (int64_t)sm_count() * 8;
            auto stage = [&](int k, bool on) {
                PipeStage st{0, 0, 0};
                if (on && k >= 0 && k < nch) {
                    st.b0 = k * pc;
                    st.nbat = B - st.b0 < pc ? B - st.b0 : pc;
                }
                return st;
            };
            for (int step = 0; step < nch + 3; ++step) {
                PipeStage zero = stage(step, true), scat = stage(step - 1, true);
                PipeStage stat = stage(step - 2, preprocess != CF_PRE_NONE), norm = stage(step - 3, preprocess != CF_PRE_NONE);
                if (zero.nbat) {
                    int64_t c = ceil_div((int64_t)zero.nbat * cells / 4, kScatterThreads * 8);
                    zero.ctas = (int)(c > 2 * cap ? 2 * cap : (c < 1 ? 1 : c));
                }
                if (scat.nbat) {
                    int64_t blocks = ceil_div(ceil_div(total * scat.nbat, B), kScatterThreads * kScatterUnroll);
                    scat.ctas = (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
                }
                if (stat.nbat) stat.ctas = stat.nbat * chunks;
                if (norm.nbat) norm.ctas = norm.nbat * chunks;
                const int grid = zero.ctas + scat.ctas + stat.ctas + norm.ctas;
                if (grid == 0) continue;
                voxel_pipeline_kernel<<<(unsigned)grid, kScatterThreads, 0, stream>>>(
                    events, offsets, B, nb, H, W, flavour, out, cells, chunk_len, chunks, hot_thr, preprocess, partials, scat, stat,
                    norm, zero);
                CF_LAUNCH_CHECK("voxel_pipeline_kernel");
            }
            return CF_OK;
        }
        const bool telescope = preprocess == CF_PRE_STD && (voxel_flags() & 32) && ws != nullptr &&
                               ws_bytes >= (size_t)B * kMaxChunks * sizeof(Partial);
        for (int b0 = 0; b0 < B; b0 += chunk) {
            const int nbat = B - b0 < chunk ? B - b0 : chunk;
            float *o = out + (size_t)b0 * cells;
            CF_CUDA(cudaMemsetAsync(o, 0, sizeof(float) * (size_t)nbat * cells, stream));
            if (total > 0) {
                // grid sized for the chunk's expected share of the events, capped at 8 CTAs per SM;
                // the kernel grid-strides, so a longer-than-average chunk is still fully processed
                int64_t blocks = ceil_div(ceil_div(total * nbat, B), kScatterThreads * kScatterUnroll);
                const int64_t cap = (int64_t)sm_count() * 8;
                if (blocks > cap) blocks = cap;
                if (blocks < 1) blocks = 1;
                if (telescope) {
                    // statistics out of the scatter's returned values (std mode; experiment, CF_VOXEL_FLAGS bit5 -- measured slower)
                    Partial *tel = reinterpret_cast<Partial *>(reinterpret_cast<char *>(ws) + (size_t)b0 * kMaxChunks * sizeof(Partial));
                    int tchunks;
                    int64_t tlen;
                    stat_geometry(nbat, cells, tchunks, tlen);
                    CF_CUDA(cudaMemsetAsync(tel, 0, sizeof(Partial) * (size_t)nbat * tchunks, stream));
                    voxel_scatter_stats_kernel<<<(unsigned)blocks, kScatterThreads, 0, stream>>>(
                        events, offsets, B, nb, H, W, flavour, out, b0, b0 + nbat, hot_thr, tel - (size_t)b0 * tchunks, tchunks);
                    CF_LAUNCH_CHECK("voxel_scatter_stats_kernel");
                } else {
                    voxel_scatter_atomic_kernel<<<(unsigned)blocks, kScatterThreads, 0, stream>>>(
                        events, offsets, B, nb, H, W, flavour, out, b0, b0 + nbat);
                    CF_LAUNCH_CHECK("voxel_scatter_atomic_kernel");
                }
            }
            if (preprocess != CF_PRE_NONE) {
                if (int rc = run_preprocess(o, o, nbat, cells, preprocess, hot_thr,
                                            reinterpret_cast<char *>(ws) + (size_t)b0 * kMaxChunks * sizeof(Partial),
                                            ws_bytes - (size_t)b0 * kMaxChunks * sizeof(Partial), stream,
                                            /*stats_done=*/telescope && total > 0)) return rc;
            }
        }
        return CF_OK;
    }
    CF_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * cells, stream));
    if (total > 0) {
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
