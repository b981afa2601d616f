// Event stream -> voxel grid, ATOMIC mode, tiled path: partition + shared-memory accumulation.
//
// Same arithmetic as voxel.cu (utils/event_process.py:15-72, 127-190, 75-123, 193-239); different data
// movement.  The L2-atomic path (voxel.cu) issues two fp32 RED per event into a grid that lives in L2;
// the B200 resolves ~70-100 G such atomics per second chip-wide, i.e. ~40 G events/s = 1.2 TB/s of
// event rows -- 18 % of the HBM rate before the grid is even written -- and then needs two more passes
// over the grid for event_preprocess.  Measured: 10-20 % of HBM with normalisation, 25-43 % without.
// Here no global atomic is issued at all:
//
//   pass A  voxel_partition_kernel   one CTA per CHUNK of <= 1024 consecutive events of one window.
//           Reads the 32-byte fp64 rows once (128-bit loads), normalises time in fp64 exactly like the
//           reference, and turns every event into a 12-byte record (cell code inside its spatial tile,
//           left weight, right weight).  A counting sort in shared memory (integer atomics) groups the
//           chunk's records by TILE (a contiguous range of P pixels x all bins of the window's grid);
//           the sorted chunk and its T+1 tile offsets are written with coalesced stores.
//   pass B  voxel_accumulate_kernel  one CTA per (window, tile), cooperative launch.  The tile
//           (nb x P cells, <= ~200 KB) lives in shared memory: zero, add the tile's run of every
//           chunk of the window (shared-memory fp32 atomics: a CAS loop in the SM, no L2 round
//           trip), reduce the tile's statistics for event_preprocess, meet the other tiles of the
//           window at a per-window arrival counter (all CTAs are co-resident: cooperative launch),
//           derive mean/std (or min/max) from the T partials in a fixed order, normalise out of
//           shared memory and write the final grid ONCE with 128-bit streaming stores.
//
// HBM traffic per event: 32 B read + 12 B written + 12 B read (the records mostly stay in L2);
// per cell: 4 B written.  Algorithmic bytes: 32 per event + 4 per cell.
// Geometry: P is chosen so that (windows per wave) x (tiles per window) fills the 148 SMs; a batch
// larger than one wave is processed in waves inside the same launch.
// Falls back to voxel.cu (return code 1) when a window's grid needs more than one CTA per SM can
// hold in shared memory (H*W > 148 * ~11 000 px at nb = 5), or B > 2048.
#include <cooperative_groups.h>

#include "voxel_common.cuh"

namespace cf {

namespace vt {
constexpr int CHUNK = 1024;            // events per pass-A CTA (4096 left too few CTAs: 15 us of serial work each)
constexpr int PART_THREADS = 256;
constexpr int PER_THREAD = CHUNK / PART_THREADS;   // 4
constexpr int ROUND = 4;               // events per thread whose loads are in flight together
constexpr int MAX_RUNS = 2048;         // chunks of one window whose run bounds pass B stages in shared memory
constexpr int MAX_B = 2048;
constexpr int MAX_T = 1024;
constexpr int ACC_THREADS = 512;
constexpr size_t TILE_BYTES_MAX = 200 * 1024;

struct Geometry {
    int ok;
    int P;            // pixels per tile
    int T;            // tiles per window
    int wpw;          // windows per wave
    int grid_b;       // CTAs of pass B
    int planes;       // nb * (2 for POL)
    int64_t max_chunks;
    size_t smem_b;    // dynamic shared memory of pass B
    // workspace layout (bytes from the start of the tiled region)
    size_t o_first, o_counters, o_offs, o_code, o_wl, o_wr, o_partials, end;
};

static Geometry geometry(int64_t total, int B, int nb, int H, int W, int flavour, int sms) {
    Geometry g{};
    const int64_t HW = (int64_t)H * W;
    g.planes = nb * (flavour == CF_FLAVOUR_POL ? 2 : 1);
    if (B < 1 || B > MAX_B || HW >= (1ll << 30)) return g;
    int64_t pmax = (int64_t)(TILE_BYTES_MAX / (sizeof(float) * g.planes)) & ~3ll;
    if (pmax < 64) return g;
    const int64_t t0 = ceil_div(HW, pmax);
    if (t0 > sms || t0 > MAX_T) return g;
    int wpw = (int)(sms / t0);
    if (wpw > B) wpw = B;
    int64_t T = sms / wpw;                       // spread each window over as many SMs as the wave allows
    if (T > MAX_T) T = MAX_T;
    int64_t P = (ceil_div(HW, T) + 3) & ~3ll;
    if (P < 64) P = 64;
    T = ceil_div(HW, P);
    g.P = (int)P;
    g.T = (int)T;
    g.wpw = wpw;
    g.grid_b = (int)(wpw * T);
    g.max_chunks = ceil_div(total > 0 ? total : 1, CHUNK) + B;
    if (g.max_chunks * CHUNK > 4 * total + (1ll << 22)) return g;   // pathological: thousands of tiny windows
    g.smem_b = (size_t)g.planes * P * sizeof(float);
    size_t o = 0;
    g.o_first = o;    o = align_up(o + (size_t)(B + 1) * sizeof(int), 256);
    g.o_counters = o; o = align_up(o + (size_t)B * sizeof(int), 256);
    g.o_offs = o;     o = align_up(o + (size_t)g.max_chunks * (T + 1) * sizeof(uint32_t), 256);
    g.o_code = o;     o = align_up(o + (size_t)g.max_chunks * CHUNK * sizeof(uint32_t), 256);
    g.o_wl = o;       o = align_up(o + (size_t)g.max_chunks * CHUNK * sizeof(float), 256);
    g.o_wr = o;       o = align_up(o + (size_t)g.max_chunks * CHUNK * sizeof(float), 256);
    g.o_partials = o; o = align_up(o + (size_t)B * T * sizeof(Partial), 256);
    g.end = o;
    g.ok = 1;
    return g;
}

// ------------------------------------------------------------------ pass A ---
__global__ void __launch_bounds__(PART_THREADS)
voxel_partition_kernel(const double *__restrict__ ev, const int64_t *__restrict__ off, int B, int nb, int H, int W,
                       int flavour, int P, int T, int *__restrict__ first_chunk, int *__restrict__ counters,
                       uint32_t *__restrict__ offs, uint32_t *__restrict__ rec_code, float *__restrict__ rec_wl,
                       float *__restrict__ rec_wr) {
    __shared__ int s_prefix[MAX_B + 1];
    __shared__ int s_hist[MAX_T + 1];
    __shared__ int s_warp[32];
    extern __shared__ __align__(16) uint32_t s_dyn[];   // sorted records of the chunk: code | wl | wr, 3 x 4 KB
    uint32_t *s_code = s_dyn;
    float *s_wl = reinterpret_cast<float *>(s_dyn + CHUNK);
    float *s_wr = reinterpret_cast<float *>(s_dyn + 2 * CHUNK);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) CF_TRACE_AT(200);
    // ---- chunks per window -> exclusive prefix (every CTA derives the same table)
    for (int b = tid; b < B; b += PART_THREADS) {
        const int64_t n = __ldg(off + b + 1) - __ldg(off + b);
        s_prefix[b + 1] = (int)((n + CHUNK - 1) / CHUNK);
    }
    for (int t = tid; t <= T; t += PART_THREADS) s_hist[t] = 0;
    if (tid == 0) s_prefix[0] = 0;
    __syncthreads();
    if (warp == 0) {  // inclusive scan of s_prefix[1..B], 32 lanes x contiguous segments
        const int per = (B + 31) / 32;
        const int s = 1 + lane * per, e = min(B + 1, s + per);
        int sum = 0;
        for (int i = s; i < e; ++i) sum += s_prefix[i];
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        int run = inc - sum;
        for (int i = s; i < e; ++i) {
            run += s_prefix[i];
            s_prefix[i] = run;
        }
    }
    __syncthreads();
    if (blockIdx.x == 0) {  // publish the table and arm the per-window arrival counters of pass B
        for (int b = tid; b <= B; b += PART_THREADS) first_chunk[b] = s_prefix[b];
        for (int b = tid; b < B; b += PART_THREADS) counters[b] = 0;
    }
    const int chunk = blockIdx.x;
    if (chunk >= s_prefix[B]) return;
    int lo = 0, hi = B - 1;  // last window w with prefix[w] <= chunk
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_prefix[mid] <= chunk) lo = mid; else hi = mid - 1;
    }
    Window w;
    w.b = lo;
    w.begin = __ldg(off + lo);
    w.end = __ldg(off + lo + 1);
    w.t0 = __ldg(ev + 4 * w.begin);
    w.span = __dsub_rn(__ldg(ev + 4 * (w.end - 1)), w.t0);
    if (w.span == 0.0) w.span = 1.0;  // event_process.py:43-44
    const int64_t first = w.begin + (int64_t)(chunk - s_prefix[lo]) * CHUNK;
    const int64_t last = min(w.end, first + CHUNK);

    if (tid == 0) CF_TRACE_AT(201);
    const int planes_per_bin = flavour == CF_FLAVOUR_POL ? 2 : 1;
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
    const size_t base = (size_t)chunk * CHUNK;
    for (int i = tid; i < n_valid; i += PART_THREADS) {
        rec_code[base + i] = s_code[i];
        rec_wl[base + i] = s_wl[i];
        rec_wr[base + i] = s_wr[i];
    }
    uint32_t *o = offs + (size_t)chunk * (T + 1);
    for (int t = tid; t <= T; t += PART_THREADS) o[t] = (uint32_t)s_hist[t];
    if (tid == 0) CF_TRACE_AT(204);
}

// ------------------------------------------------------------------ pass B ---
__global__ void __launch_bounds__(ACC_THREADS, 1)
voxel_accumulate_kernel(const uint32_t *__restrict__ rec_code, const float *__restrict__ rec_wl,
                        const float *__restrict__ rec_wr, const uint32_t *__restrict__ offs,
                        const int *__restrict__ first_chunk, int *counters, Partial *partials, int B, int planes,
                        int right_stride /* cells between a bin and the next one inside the tile */, int64_t HW, int P,
                        int T, int wpw, int preprocess, float hot_thr, float *__restrict__ out) {
    extern __shared__ __align__(16) float tile[];   // [planes][P]
    __shared__ Partial s_part[ACC_THREADS / 32];
    __shared__ int s_lo[MAX_RUNS], s_hi[MAX_RUNS];
    __shared__ double s_a, s_inv;
    __shared__ int s_identity;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = blockIdx.x % T, slot = blockIdx.x / T;
    const int cells = planes * P;
    const int px0 = t * P;
    const int valid = (int)min((int64_t)P, HW - px0);   // pixels of this tile inside the grid (> 0 by construction)

    int wave = 0;
    for (int b = slot; b < B; b += wpw, ++wave) {
        if (tid == 0) CF_TRACE_AT(8 * wave + 0);
        // ---- zero the tile
        for (int i = tid; i < cells / 4; i += ACC_THREADS) reinterpret_cast<float4 *>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        // ---- accumulate this tile's run of every chunk of the window.  Run bounds first (one round of
        //      independent loads, staged in shared memory), then warp <-> 4 runs at a time so that the
        //      record loads of 4 runs are in flight together.
        const int c0 = __ldg(first_chunk + b), c1 = __ldg(first_chunk + b + 1);
        for (int cb = c0; cb < c1; cb += MAX_RUNS) {
            const int nrun = min(MAX_RUNS, c1 - cb);
            for (int i = tid; i < nrun; i += ACC_THREADS) {
                const uint32_t *o = offs + (size_t)(cb + i) * (T + 1) + t;
                s_lo[i] = (int)__ldg(o);
                s_hi[i] = (int)__ldg(o + 1);
            }
            __syncthreads();
            if (tid == 0) CF_TRACE_AT(8 * wave + 1);
            constexpr int U = 4;
            for (int r0 = warp * U; r0 < nrun; r0 += (ACC_THREADS / 32) * U) {
                uint32_t code[U];
                float wa[U], wb[U];
                bool ok[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int r = r0 + u;
                    const int lo = r < nrun ? s_lo[r] : 0, hi = r < nrun ? s_hi[r] : 0;
                    ok[u] = lo + lane < hi;
                    const size_t at = (size_t)(cb + r) * CHUNK + lo + lane;
                    if (ok[u]) { code[u] = __ldg(rec_code + at); wa[u] = __ldg(rec_wl + at); wb[u] = __ldg(rec_wr + at); }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (ok[u]) {
                        const int cell = (int)(code[u] & 0x7fffffffu);
                        atomicAdd(tile + cell, wa[u]);
                        if (!(code[u] >> 31)) atomicAdd(tile + cell + right_stride, wb[u]);
                    }
                }
#pragma unroll 1
                for (int u = 0; u < U; ++u) {  // runs longer than one warp (hot tiles): the rest, 32 at a time
                    const int r = r0 + u;
                    if (r >= nrun) break;
                    const size_t base = (size_t)(cb + r) * CHUNK;
                    for (int i = s_lo[r] + 32 + lane; i < s_hi[r]; i += 32) {
                        const uint32_t cd = __ldg(rec_code + base + i);
                        const float x = __ldg(rec_wl + base + i), y = __ldg(rec_wr + base + i);
                        const int cell = (int)(cd & 0x7fffffffu);
                        atomicAdd(tile + cell, x);
                        if (!(cd >> 31)) atomicAdd(tile + cell + right_stride, y);
                    }
                }
            }
            __syncthreads();
        }

        if (tid == 0) CF_TRACE_AT(8 * wave + 2);
        double a = 0.0, inv = 1.0;
        bool identity = true;
        if (preprocess != CF_PRE_NONE) {
            // ---- statistics of the tile's valid cells (hot pixels filtered first, event_process.py:196-198)
            // per plane and thread a short fp32 partial (<= P/512 ~ 20 terms), promoted to fp64 across planes,
            // threads and tiles: an fp64 add + fma per CELL cost 2 us per tile on the fp64 pipe
            double sum = 0.0, sumsq = 0.0;
            long long nnz = 0;
            float mn = INFINITY, mx = -INFINITY;
            for (int k = 0; k < planes; ++k) {
                const float *row = tile + k * P;
                float fs = 0.f, fq = 0.f;
                int fn = 0;
                for (int j = tid; j < valid; j += ACC_THREADS) {
                    const float v = hot_filter(row[j], hot_thr);
                    fs += v;
                    fq = fmaf(v, v, fq);
                    fn += (v != 0.f);
                    mn = fminf(mn, v);
                    mx = fmaxf(mx, v);
                }
                sum += (double)fs;
                sumsq += (double)fq;
                nnz += fn;
            }
            sum = warp_sum(sum); sumsq = warp_sum(sumsq); nnz = warp_sum(nnz);
            mn = warp_min(mn); mx = warp_max(mx);
            if (lane == 0) s_part[warp] = Partial{sum, sumsq, nnz, mn, mx};
            __syncthreads();
            if (tid == 0) {
                Partial p = s_part[0];
                for (int k = 1; k < ACC_THREADS / 32; ++k) {
                    p.sum += s_part[k].sum; p.sumsq += s_part[k].sumsq; p.nnz += s_part[k].nnz;
                    p.mn = fminf(p.mn, s_part[k].mn); p.mx = fmaxf(p.mx, s_part[k].mx);
                }
                partials[(size_t)b * T + t] = p;
                CF_TRACE_AT(8 * wave + 3);
                __threadfence();
                atomicAdd(counters + b, 1);
                // all T CTAs of window b are resident (cooperative launch, same wave): bounded spin
                unsigned spins = 0;
                while (*reinterpret_cast<volatile int *>(counters + b) < T) {
                    __nanosleep(64);
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
                    if (preprocess == CF_PRE_STD) {
                        s_identity = tn == 0;  // event_process.py:205 -- untouched when there is no non-zero entry
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
            if (tid == 0) CF_TRACE_AT(8 * wave + 4);
        }
        // ---- normalise out of shared memory, write the final grid once
        auto norm = [&](float raw) -> float {
            if (preprocess == CF_PRE_NONE) return raw;
            const float v = hot_filter(raw, hot_thr);
            if (identity) return v;
            if (preprocess == CF_PRE_STD) return (v != 0.f) ? (float)(((double)v - a) * inv) : 0.f;
            return (float)(((double)v - a) * inv);
        };
        float *ob = out + (size_t)b * planes * HW + px0;
        if ((HW & 3) == 0) {   // plane starts and tile starts are 16-byte aligned
            const int v4 = valid / 4;   // valid % 4 == 0: HW % 4 == 0 and P % 4 == 0
            for (int idx = tid; idx < planes * v4; idx += ACC_THREADS) {
                const int k = idx / v4, j = idx - k * v4;
                const float4 q = reinterpret_cast<const float4 *>(tile + k * P)[j];
                st_cs4(reinterpret_cast<float4 *>(ob + (size_t)k * HW) + j, make_float4(norm(q.x), norm(q.y), norm(q.z), norm(q.w)));
            }
        } else {
            for (int idx = tid; idx < planes * valid; idx += ACC_THREADS) {
                const int k = idx / valid, j = idx - k * valid;
                st_cs(ob + (size_t)k * HW + j, norm(tile[k * P + j]));
            }
        }
        __syncthreads();   // the tile is re-zeroed by the next wave
        if (tid == 0) CF_TRACE_AT(8 * wave + 5);
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
    int *first_chunk = reinterpret_cast<int *>(w8 + g.o_first);
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
        CF_CUDA(cudaFuncSetAttribute(voxel_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_BYTES_MAX + 1024));
        opt_in[dev & 63] = true;
    }
    voxel_partition_kernel<<<(unsigned)g.max_chunks, PART_THREADS, kPartSmem, stream>>>(
        events, offsets, B, nb, H, W, flavour, g.P, g.T, first_chunk, counters, offs, rec_code, rec_wl, rec_wr);
    CF_LAUNCH_CHECK("voxel_partition_kernel");

    const int planes = g.planes;
    const int right_stride = (flavour == CF_FLAVOUR_POL ? 2 : 1) * g.P;
    const int64_t HW = (int64_t)H * W;
    int P = g.P, T = g.T, wpw = g.wpw;
    const uint32_t *c_code = rec_code;
    const float *c_wl = rec_wl, *c_wr = rec_wr;
    const uint32_t *c_offs = offs;
    const int *c_first = first_chunk;
    int Bv = B, pre = preprocess, rs = right_stride, pl = planes;
    int64_t hw = HW;
    float thr = hot_thr;
    void *args[] = {&c_code, &c_wl, &c_wr, &c_offs, &c_first, &counters, &partials, &Bv, &pl, &rs, &hw, &P, &T, &wpw, &pre, &thr, &out};
    cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(voxel_accumulate_kernel), dim3((unsigned)g.grid_b),
                                                dim3(ACC_THREADS), args, g.smem_b, stream);
    count_launch("voxel_accumulate_kernel");
    if (e != cudaSuccess) {
        set_error("cooperative launch of voxel_accumulate_kernel (%d CTAs, %zu B smem) failed: %s", g.grid_b, g.smem_b,
                  cudaGetErrorString(e));
        return CF_ERR_CUDA;
    }
    return CF_OK;
}

}  // namespace cf
