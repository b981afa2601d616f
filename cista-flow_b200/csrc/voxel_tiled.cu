// Event stream -> voxel grid, ATOMIC mode, tiled path: partition + shared-memory accumulation.
//
// Same arithmetic as voxel.cu (utils/event_process.py:15-72, 127-190, 75-123, 193-239); different data
// movement.  The L2-atomic path (voxel.cu) zero-fills the grid, issues two fp32 RED per event into it, then
// reads it for the statistics and reads + writes it again for event_preprocess: four passes over a grid that
// the algorithm writes once, plus ~105 G RED/s chip-wide (profiles/r02/atomics_l2_probe.txt) -- 25 % of HBM.
// Here the grid is written exactly once and no global atomic is issued:
//
//   pass A  voxel_partition_kernel   CTA (x, window): CHUNKs of 2048 consecutive events of one window.
//           Reads the 32-byte fp64 rows once (128-bit loads), normalises time in fp64 exactly like the
//           reference, and turns every event into a 12-byte record (cell code inside its spatial tile,
//           left weight, right weight).  A counting sort in shared memory (integer atomics) groups the
//           chunk's records by TILE (a contiguous range of P pixels x all bins of the window's grid);
//           the sorted chunk lands at the chunk's own event positions (coalesced), its T+1 tile offsets
//           in a small table.  No prefix over the windows: chunk slot = begin / CHUNK + window + chunk.
//   pass B  voxel_tile_kernel        persistent, cooperative launch, 2-4 CTAs per SM; item = (window, tile).
//           The tile (nb x P cells) lives in shared memory: zero, add the tile's run of every chunk of the
//           window with a shared-memory CAS loop (221 G adds/s chip-wide against 105 G/s for L2 REDs, same
//           probe).  The CAS loop hands back each add's OLD value for free, so the window statistics of
//           event_preprocess come out of the scatter itself: sum f(new) - f(old) telescopes per cell to the
//           final grid's sum / sum of squares / non-zero count (f = hot-pixel filter) -- no statistics pass.
//           (Round 1 tried the same with returning L2 atomics: ATOMG is 4x slower than RED.  In shared memory
//           the old value is a by-product.)  The tiles of a window meet at a per-window arrival counter,
//           every CTA derives mean/std from the T partials in the same fixed order, normalises out of shared
//           memory (fp32 map, the arithmetic of voxel_normalise_kernel) and writes the grid ONCE with
//           128-bit streaming stores.  While one CTA of an SM waits or loads, the other one zeroes, adds or
//           writes.  Deadlock-free: items are ordered by window and T <= grid size, so a CTA holds at most
//           one item per window and every CTA a waiter depends on is at an earlier window (induction).
//
// HBM traffic per event: 32 B read + 12 B written + 12 B read (the records mostly stay in L2);
// per cell: 4 B written.  Algorithmic bytes: 32 per event + 4 per cell.
// Falls back to voxel.cu (return code 1) when a window needs more tiles than the launch has CTAs
// (H*W > ~296 * 5 600 px at nb = 5), B > 65535, or for the MVSEC flavour.
#include <cooperative_groups.h>

#include "voxel_common.cuh"

namespace cf {

namespace vt {
constexpr int CHUNK = 2048;            // events per pass-A work item
constexpr int PART_THREADS = 512;
constexpr int PER_THREAD = CHUNK / PART_THREADS;   // 4
constexpr int ROUND = 4;               // events per thread whose loads are in flight together
constexpr int MAX_T = 1024;
constexpr int ACC_THREADS = 512;
constexpr int ACC_WARPS = ACC_THREADS / 32;
constexpr size_t TILE_BYTES_MAX = 108 * 1024;   // two CTAs per SM

struct Geometry {
    int ok;
    int P;            // pixels per tile
    int T;            // tiles per window
    int grid_b;       // CTAs of pass B
    int grid_ax;      // pass A: CTAs per window
    int planes;       // nb * (2 for POL)
    int64_t max_slots;
    size_t smem_b;    // dynamic shared memory of pass B
    // workspace layout (bytes from the start of the tiled region)
    size_t o_counters, o_offs, o_code, o_wl, o_wr, o_partials, end;
};

static Geometry geometry(int64_t total, int B, int nb, int H, int W, int flavour, int sms) {
    Geometry g{};
    const int64_t HW = (int64_t)H * W;
    g.planes = nb * (flavour == CF_FLAVOUR_POL ? 2 : 1);
    if (B < 1 || B > 65535 || HW >= (1ll << 28) || total < 0) return g;
    const int64_t pmax = (int64_t)(TILE_BYTES_MAX / (sizeof(float) * g.planes)) & ~3ll;
    if (pmax < 64) return g;
    const int64_t t0 = ceil_div(HW, pmax);
    // small batches: more (smaller) tiles so that every SM gets work, but runs of >= 16 records per chunk and tile
    int64_t want = ceil_div(2 * (int64_t)sms, B);
    if (want > CHUNK / 16) want = CHUNK / 16;
    if (want > HW / 256) want = HW / 256;
    int64_t T = t0 > want ? t0 : want;
    if (T < 1) T = 1;
    int64_t P = (ceil_div(HW, T) + 3) & ~3ll;
    T = ceil_div(HW, P);
    if (T > MAX_T) return g;
    g.P = (int)P;
    g.T = (int)T;
    g.smem_b = (size_t)g.planes * P * sizeof(float);
    int per_sm = (int)((227 * 1024) / (g.smem_b + 2048));
    if (per_sm > 2) per_sm = 2;                  // __launch_bounds__(512, 2): the register file holds two CTAs
    if (per_sm < 1) return g;
    const int64_t cap = (int64_t)sms * per_sm;
    if (T > cap) return g;                       // a CTA must never hold two items of one window (see header)
    const int64_t items = (int64_t)B * T;
    g.grid_b = (int)(items < cap ? items : cap);
    int64_t ax = ceil_div(ceil_div(total > 0 ? total : 1, B), CHUNK);
    if (ax < 1) ax = 1;
    if (ax > 4096) ax = 4096;                    // longer windows loop
    g.grid_ax = (int)ax;
    g.max_slots = (total > 0 ? total : 1) / CHUNK + B + 1;
    size_t o = 0;
    g.o_counters = o; o = align_up(o + (size_t)B * sizeof(int), 256);
    g.o_offs = o;     o = align_up(o + (size_t)g.max_slots * (T + 1) * sizeof(uint32_t), 256);
    g.o_code = o;     o = align_up(o + (size_t)(total > 0 ? total : 1) * sizeof(uint32_t), 256);
    g.o_wl = o;       o = align_up(o + (size_t)(total > 0 ? total : 1) * sizeof(float), 256);
    g.o_wr = o;       o = align_up(o + (size_t)(total > 0 ? total : 1) * sizeof(float), 256);
    g.o_partials = o; o = align_up(o + (size_t)B * T * sizeof(Partial), 256);
    g.end = o;
    g.ok = 1;
    return g;
}

// ------------------------------------------------------------------ pass A ---
__global__ void __launch_bounds__(PART_THREADS, 2)
voxel_partition_kernel(const double *__restrict__ ev, const int64_t *__restrict__ off, int nb, int H, int W,
                       int flavour, int P, int T, int *__restrict__ counters, uint32_t *__restrict__ offs,
                       uint32_t *__restrict__ rec_code, float *__restrict__ rec_wl, float *__restrict__ rec_wr) {
    __shared__ int s_hist[MAX_T + 1];
    __shared__ int s_warp[PART_THREADS / 32];
    extern __shared__ __align__(16) uint32_t s_dyn[];   // sorted records of the chunk: code | wl | wr, 3 x 8 KB
    uint32_t *s_code = s_dyn;
    float *s_wl = reinterpret_cast<float *>(s_dyn + CHUNK);
    float *s_wr = reinterpret_cast<float *>(s_dyn + 2 * CHUNK);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;

    if (tid == 0) CF_TRACE_AT(200);
    Window w;
    w.b = b;
    w.begin = __ldg(off + b);
    w.end = __ldg(off + b + 1);
    if (blockIdx.x == 0 && tid == 0) counters[b] = 0;   // arms the per-window arrival counter of pass B
    const int64_t n = w.end - w.begin;
    if (n <= 0 || (int64_t)blockIdx.x * CHUNK >= n) return;
    w.t0 = __ldg(ev + 4 * w.begin);
    w.span = __dsub_rn(__ldg(ev + 4 * (w.end - 1)), w.t0);
    if (w.span == 0.0) w.span = 1.0;  // event_process.py:43-44
    const int64_t slot0 = w.begin / CHUNK + b;
    const int planes_per_bin = flavour == CF_FLAVOUR_POL ? 2 : 1;

    for (int64_t c = blockIdx.x; c * CHUNK < n; c += gridDim.x) {
        const int64_t first = w.begin + c * CHUNK;
        const int64_t last = min(w.end, first + CHUNK);
        for (int t = tid; t <= T; t += PART_THREADS) s_hist[t] = 0;
        __syncthreads();
        if (tid == 0) CF_TRACE_AT(201);
        uint32_t code[PER_THREAD], where[PER_THREAD];   // where = tile | rank << 16, 0xffffffff: dropped
        float wl[PER_THREAD], wr[PER_THREAD];
#pragma unroll
        for (int r = 0; r < PER_THREAD; r += ROUND) {
            Event e[ROUND];
#pragma unroll
            for (int k = 0; k < ROUND; ++k) {  // all loads of the round in flight first
                const int64_t i = first + (int64_t)(r + k) * PART_THREADS + tid;
                if (i < last) e[k] = load_event(ev, i);
            }
#pragma unroll
            for (int k = 0; k < ROUND; ++k) {
                const int64_t i = first + (int64_t)(r + k) * PART_THREADS + tid;
                where[r + k] = 0xffffffffu;
                if (i >= last) continue;
                const Binned bb = bin_event(e[k], w, nb, H, W, flavour);
                if (!bb.ok) continue;
                if (flavour == CF_FLAVOUR_TORCH) {
                    weights_f32(bb, wl[r + k], wr[r + k]);
                } else {
                    double dl, dr;
                    weights_f64(bb, dl, dr);
                    wl[r + k] = (float)dl;
                    wr[r + k] = (float)dr;
                }
                const int pix = bb.y * W + bb.x;
                const int tile = pix / P;
                const int local = pix - tile * P;
                // cell index inside the tile's shared-memory image [planes][P]; bit 31: no right neighbour
                code[r + k] = (uint32_t)((bb.bin * planes_per_bin + bb.chan) * P + local) | (bb.bin + 1 < nb ? 0u : 0x80000000u);
                const int rank = atomicAdd(&s_hist[tile], 1);  // integer, shared memory: order-independent totals
                where[r + k] = (uint32_t)tile | ((uint32_t)rank << 16);
            }
        }
        __syncthreads();
        if (tid == 0) CF_TRACE_AT(202);
        // ---- exclusive scan of the T tile counts (in place; s_hist[T] = records of the chunk)
        {
            const int per = (T + PART_THREADS) / PART_THREADS;   // covers T + 1 entries
            const int s = tid * per, e = min(T + 1, s + per);
            int sum = 0;
            for (int i = s; i < e; ++i) sum += s_hist[i];
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (lane == 31) s_warp[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                const int tot = lane < PART_THREADS / 32 ? s_warp[lane] : 0;
                int ti = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, ti, o);
                    if (lane >= o) ti += v;
                }
                if (lane < PART_THREADS / 32) s_warp[lane] = ti - tot;
            }
            __syncthreads();
            int run = s_warp[warp] + inc - sum;
            for (int i = s; i < e; ++i) {
                const int v = s_hist[i];
                s_hist[i] = run;
                run += v;
            }
        }
        __syncthreads();
        // ---- records to their sorted slot in shared memory, then coalesced to global
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k) {
            if (where[k] != 0xffffffffu) {
                const int pos = s_hist[where[k] & 0xffffu] + (int)(where[k] >> 16);
                s_code[pos] = code[k];
                s_wl[pos] = wl[k];
                s_wr[pos] = wr[k];
            }
        }
        __syncthreads();
        if (tid == 0) CF_TRACE_AT(203);
        const int n_valid = s_hist[T];
        for (int i = tid; i < n_valid; i += PART_THREADS) {
            rec_code[first + i] = s_code[i];
            rec_wl[first + i] = s_wl[i];
            rec_wr[first + i] = s_wr[i];
        }
        uint32_t *o = offs + (size_t)(slot0 + c) * (T + 1);
        for (int t = tid; t <= T; t += PART_THREADS) o[t] = (uint32_t)s_hist[t];
        __syncthreads();   // the shared arrays are reused by the next chunk
        if (tid == 0) CF_TRACE_AT(204);
    }
}

// ------------------------------------------------------------------ pass B ---
// fp32 add into the tile; returns nothing but feeds the telescoping statistics with (old, new)
template <bool STATS>
__device__ __forceinline__ void tile_add(float *cell, float w, float hot_thr, double &d_sum, double &d_sq, int &d_nnz) {
    unsigned *a = reinterpret_cast<unsigned *>(cell);
    unsigned old = *a, assumed;
    float nw;
    do {
        assumed = old;
        nw = __fadd_rn(__uint_as_float(assumed), w);
        old = atomicCAS(a, assumed, __float_as_uint(nw));
    } while (old != assumed);
    if (STATS) {
        const float fo = hot_filter(__uint_as_float(old), hot_thr), fn = hot_filter(nw, hot_thr);
        d_sum += (double)fn - (double)fo;
        d_sq += (double)fn * (double)fn - (double)fo * (double)fo;   // squares of fp32 values are exact in fp64
        d_nnz += (int)(fn != 0.f) - (int)(fo != 0.f);
    }
}

template <int PRE>
__global__ void __launch_bounds__(ACC_THREADS, 2)
voxel_tile_kernel(const uint32_t *__restrict__ rec_code, const float *__restrict__ rec_wl,
                  const float *__restrict__ rec_wr, const uint32_t *__restrict__ offs,
                  const int64_t *__restrict__ off, int *counters, Partial *partials, int B, int planes,
                  int right_stride /* cells between a bin and the next one inside the tile */, int64_t HW, int P,
                  int T, float hot_thr, float *__restrict__ out) {
    extern __shared__ __align__(16) float tile[];   // [planes][P]
    __shared__ Partial s_part[ACC_WARPS];
    __shared__ float s_a, s_b;
    __shared__ int s_identity;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cells = planes * P;
    const int64_t items = (int64_t)B * T;

    int round = 0;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x, ++round) {
        const int b = (int)(item / T), t = (int)(item - (int64_t)b * T);
        const int px0 = t * P;
        const int valid = (int)min((int64_t)P, HW - px0);   // pixels of this tile inside the grid (> 0 by construction)
        if (tid == 0) CF_TRACE_AT(8 * round + 0);
        // ---- zero the tile
        for (int i = tid; i < cells / 4; i += ACC_THREADS) reinterpret_cast<float4 *>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t begin = __ldg(off + b), end = __ldg(off + b + 1);
        const int nchunks = (int)((end - begin + CHUNK - 1) / CHUNK);
        const int64_t slot0 = begin / CHUNK + b;
        __syncthreads();
        // ---- accumulate this tile's run of every chunk of the window: warp <-> 4 chunks at a time, so that the
        //      run bounds of 4 chunks, then their records, are in flight together
        double d_sum = 0.0, d_sq = 0.0;
        int d_nnz = 0;
        constexpr int U = 4;
        for (int c0 = warp * U; c0 < nchunks; c0 += ACC_WARPS * U) {
            int lo[U], hi[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int c = c0 + u;
                lo[u] = hi[u] = 0;
                if (c < nchunks) {
                    const uint32_t *o = offs + (size_t)(slot0 + c) * (T + 1) + t;
                    lo[u] = (int)__ldg(o);
                    hi[u] = (int)__ldg(o + 1);
                }
            }
            uint32_t code[U];
            float wa[U], wb[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                ok[u] = lo[u] + lane < hi[u];
                const size_t at = (size_t)(begin + (int64_t)(c0 + u) * CHUNK) + lo[u] + lane;
                if (ok[u]) { code[u] = __ldg(rec_code + at); wa[u] = __ldg(rec_wl + at); wb[u] = __ldg(rec_wr + at); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (ok[u]) {
                    const int cell = (int)(code[u] & 0x7fffffffu);
                    tile_add<PRE == CF_PRE_STD>(tile + cell, wa[u], hot_thr, d_sum, d_sq, d_nnz);
                    if (!(code[u] >> 31)) tile_add<PRE == CF_PRE_STD>(tile + cell + right_stride, wb[u], hot_thr, d_sum, d_sq, d_nnz);
                }
            }
#pragma unroll 1
            for (int u = 0; u < U; ++u) {  // runs longer than one warp (hot tiles): the rest, 32 at a time
                if (c0 + u >= nchunks) break;
                const size_t base = (size_t)(begin + (int64_t)(c0 + u) * CHUNK);
                for (int i = lo[u] + 32 + lane; i < hi[u]; i += 32) {
                    const uint32_t cd = __ldg(rec_code + base + i);
                    const float x = __ldg(rec_wl + base + i), y = __ldg(rec_wr + base + i);
                    const int cell = (int)(cd & 0x7fffffffu);
                    tile_add<PRE == CF_PRE_STD>(tile + cell, x, hot_thr, d_sum, d_sq, d_nnz);
                    if (!(cd >> 31)) tile_add<PRE == CF_PRE_STD>(tile + cell + right_stride, y, hot_thr, d_sum, d_sq, d_nnz);
                }
            }
        }
        __syncthreads();
        if (tid == 0) CF_TRACE_AT(8 * round + 1);

        float a = 0.f, inv = 1.f;
        bool identity = true;
        if (PRE != CF_PRE_NONE) {
            // ---- this tile's share of the window statistics (hot pixels filtered first, event_process.py:196-198)
            double sum = 0.0, sumsq = 0.0;
            long long nnz = 0;
            float mn = INFINITY, mx = -INFINITY;
            if (PRE == CF_PRE_STD) {          // telescoped out of the scatter
                sum = d_sum; sumsq = d_sq; nnz = d_nnz;
            } else {                          // min / max do not telescope: one pass over the tile's valid cells
                for (int k = 0; k < planes; ++k) {
                    const float *row = tile + k * P;
                    for (int j = tid; j < valid; j += ACC_THREADS) {
                        const float v = hot_filter(row[j], hot_thr);
                        mn = fminf(mn, v);
                        mx = fmaxf(mx, v);
                    }
                }
            }
            sum = warp_sum(sum); sumsq = warp_sum(sumsq); nnz = warp_sum(nnz);
            mn = warp_min(mn); mx = warp_max(mx);
            if (lane == 0) s_part[warp] = Partial{sum, sumsq, nnz, mn, mx};
            __syncthreads();
            if (tid == 0) {
                Partial p = s_part[0];
                for (int k = 1; k < ACC_WARPS; ++k) {
                    p.sum += s_part[k].sum; p.sumsq += s_part[k].sumsq; p.nnz += s_part[k].nnz;
                    p.mn = fminf(p.mn, s_part[k].mn); p.mx = fmaxf(p.mx, s_part[k].mx);
                }
                partials[(size_t)b * T + t] = p;
                CF_TRACE_AT(8 * round + 2);
                __threadfence();
                atomicAdd(counters + b, 1);
                // every tile of window b is held by a CTA of this (cooperative) launch that reaches this point
                // without waiting on a later window: bounded spin
                unsigned spins = 0;
                while (*reinterpret_cast<volatile int *>(counters + b) < T) {
                    __nanosleep(32);
                    if (++spins > (1u << 24)) __trap();
                }
                __threadfence();
            }
            __syncthreads();
            if (warp == 0) {  // every CTA of the window combines the T partials in the same fixed order
                const Partial *p = partials + (size_t)b * T;
                double ts = 0.0, tq = 0.0;
                long long tn = 0;
                float tmn = INFINITY, tmx = -INFINITY;
                for (int k = lane; k < T; k += 32) {
                    const volatile Partial *q = p + k;   // written by other CTAs of this launch: bypass L1
                    ts += q->sum; tq += q->sumsq; tn += q->nnz;
                    tmn = fminf(tmn, q->mn); tmx = fmaxf(tmx, q->mx);
                }
                ts = warp_sum(ts); tq = warp_sum(tq); tn = warp_sum(tn);
                tmn = warp_min(tmn); tmx = warp_max(tmx);
                if (lane == 0) {
                    if (PRE == CF_PRE_STD) {
                        s_identity = tn == 0;  // event_process.py:205 -- untouched when there is no non-zero entry
                        const double mean = tn ? ts / (double)tn : 0.0;
                        const double var = tn ? tq / (double)tn - mean * mean : 0.0;
                        s_a = (float)mean;
                        s_b = (float)(1.0 / (sqrt(fmax(var, 0.0)) + 1e-8));
                    } else {
                        s_identity = 0;
                        s_a = tmn;
                        s_b = (float)(1.0 / ((double)tmx - (double)tmn + 1e-8));
                    }
                }
            }
            __syncthreads();
            a = s_a; inv = s_b; identity = s_identity != 0;
            if (tid == 0) CF_TRACE_AT(8 * round + 3);
        }
        // ---- normalise out of shared memory (the fp32 map of voxel_normalise_kernel), write the final grid once
        auto norm = [&](float raw) -> float {
            if (PRE == CF_PRE_NONE) return raw;
            const float v = hot_filter(raw, hot_thr);
            if (identity) return v;
            const float r = (v - a) * inv;
            return (PRE == CF_PRE_STD && v == 0.f) ? 0.f : r;
        };
        float *ob = out + (size_t)b * planes * HW + px0;
        if ((HW & 3) == 0) {   // plane starts and tile starts are 16-byte aligned
            const int v4 = valid / 4;   // valid % 4 == 0: HW % 4 == 0 and P % 4 == 0
            for (int k = 0; k < planes; ++k) {
                const float4 *src = reinterpret_cast<const float4 *>(tile + k * P);
                float4 *dst = reinterpret_cast<float4 *>(ob + (size_t)k * HW);
                for (int j = tid; j < v4; j += ACC_THREADS) {
                    const float4 q = src[j];
                    st_cs4(dst + j, make_float4(norm(q.x), norm(q.y), norm(q.z), norm(q.w)));
                }
            }
        } else {
            for (int k = 0; k < planes; ++k)
                for (int j = tid; j < valid; j += ACC_THREADS) st_cs(ob + (size_t)k * HW + j, norm(tile[k * P + j]));
        }
        __syncthreads();   // the tile is re-zeroed by the next item
        if (tid == 0) CF_TRACE_AT(8 * round + 4);
    }
}
}  // namespace vt

CF_DEFINE_TRACE_SETTER(cf_trace_buffer_voxel)

size_t voxel_tiled_workspace_bytes(int64_t total, int B, int nb, int H, int W, int flavour) {
    const vt::Geometry g = vt::geometry(total, B, nb, H, W, flavour, sm_count());
    return g.ok ? g.end : 0;
}

// CF_OK / error, or 1 when the tiled path does not apply (caller uses the L2-atomic path of voxel.cu)
int launch_voxel_tiled(const double *events, const int64_t *offsets, int64_t total, int B, int nb, int H, int W,
                       int flavour, int preprocess, float hot_thr, float *out, void *ws, size_t ws_bytes,
                       cudaStream_t stream) {
    using namespace vt;
    const Geometry g = geometry(total, B, nb, H, W, flavour, sm_count());
    if (!g.ok || total <= 0) return 1;
    CF_REQUIRE(ws && ws_bytes >= g.end, CF_ERR_WORKSPACE, "cf_voxel_bin: workspace too small (%zu < %zu)", ws_bytes, g.end);
    CF_REQUIRE(aligned16(ws), CF_ERR_ALIGN, "cf_voxel_bin: workspace not 16-byte aligned");
    char *w8 = reinterpret_cast<char *>(ws);
    int *counters = reinterpret_cast<int *>(w8 + g.o_counters);
    uint32_t *offs = reinterpret_cast<uint32_t *>(w8 + g.o_offs);
    uint32_t *rec_code = reinterpret_cast<uint32_t *>(w8 + g.o_code);
    float *rec_wl = reinterpret_cast<float *>(w8 + g.o_wl);
    float *rec_wr = reinterpret_cast<float *>(w8 + g.o_wr);
    Partial *partials = reinterpret_cast<Partial *>(w8 + g.o_partials);

    int dev = 0;
    CF_CUDA(cudaGetDevice(&dev));
    static bool opt_in[64] = {};
    constexpr size_t kPartSmem = 3 * CHUNK * sizeof(uint32_t);
    if (!opt_in[dev & 63]) {
        CF_CUDA(cudaFuncSetAttribute(voxel_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPartSmem));
        CF_CUDA(cudaFuncSetAttribute(voxel_tile_kernel<CF_PRE_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_BYTES_MAX + 1024));
        CF_CUDA(cudaFuncSetAttribute(voxel_tile_kernel<CF_PRE_STD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_BYTES_MAX + 1024));
        CF_CUDA(cudaFuncSetAttribute(voxel_tile_kernel<CF_PRE_MAXMIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_BYTES_MAX + 1024));
        opt_in[dev & 63] = true;
    }
    voxel_partition_kernel<<<dim3((unsigned)g.grid_ax, (unsigned)B), PART_THREADS, kPartSmem, stream>>>(
        events, offsets, nb, H, W, flavour, g.P, g.T, counters, offs, rec_code, rec_wl, rec_wr);
    CF_LAUNCH_CHECK("voxel_partition_kernel");

    const int planes = g.planes;
    const int right_stride = (flavour == CF_FLAVOUR_POL ? 2 : 1) * g.P;
    const int64_t HW = (int64_t)H * W;
    int P = g.P, T = g.T;
    const uint32_t *c_code = rec_code;
    const float *c_wl = rec_wl, *c_wr = rec_wr;
    const uint32_t *c_offs = offs;
    int Bv = B, rs = right_stride, pl = planes;
    int64_t hw = HW;
    float thr = hot_thr;
    void *args[] = {&c_code, &c_wl, &c_wr, &c_offs, &offsets, &counters, &partials, &Bv, &pl, &rs, &hw, &P, &T, &thr, &out};
    const void *fn = preprocess == CF_PRE_STD ? reinterpret_cast<const void *>(voxel_tile_kernel<CF_PRE_STD>)
                     : preprocess == CF_PRE_MAXMIN ? reinterpret_cast<const void *>(voxel_tile_kernel<CF_PRE_MAXMIN>)
                                                   : reinterpret_cast<const void *>(voxel_tile_kernel<CF_PRE_NONE>);
    int resident = 0;
    CF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, fn, ACC_THREADS, g.smem_b));
    const int64_t cap = (int64_t)resident * sm_count();
    if (cap < g.T) return 1;   // cannot keep a whole window co-resident: the L2-atomic path takes over
    const int grid_b = (int)(g.grid_b < cap ? g.grid_b : cap);
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid_b), dim3(ACC_THREADS), args, g.smem_b, stream);
    count_launch("voxel_tile_kernel");
    if (e != cudaSuccess) {
        set_error("cooperative launch of voxel_tile_kernel (%d CTAs, %zu B smem) failed: %s", g.grid_b, g.smem_b,
                  cudaGetErrorString(e));
        return CF_ERR_CUDA;
    }
    return CF_OK;
}

}  // namespace cf
