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

#include "tma.cuh"
#include "voxel_common.cuh"

namespace cf {

namespace vt {
constexpr int CHUNK = 2048;            // events per pass-A work item
constexpr int PART_THREADS = 512;
constexpr int PER_THREAD = CHUNK / PART_THREADS;   // 4
constexpr int MAX_T = 1024;
constexpr int ACC_THREADS = 512;       // threads per warp group of pass B (ACC + FIN = 1024 per CTA)
constexpr int ACC_WARPS = ACC_THREADS / 32;
constexpr size_t TILE_BYTES_MAX = 108 * 1024;   // two tile buffers per CTA, one CTA per SM
constexpr size_t PART_SMEM = (size_t)CHUNK * 32 /*raw rows*/ + 3 * (size_t)CHUNK * 4 /*sorted records*/;   // 88 KB: 2 CTAs per SM

struct alignas(16) WindowInfo {
    int64_t begin, end;
    double t0, span;
};

struct Geometry {
    int ok;
    int P;            // pixels per tile
    int T;            // tiles per window
    int grid_b;       // CTAs of pass B
    int max_c;        // pass A: chunk columns of the virtual [B][max_c] item grid (longer windows wrap around)
    int grid_a;       // CTAs of pass A
    int planes;       // nb * (2 for POL)
    int64_t max_slots;
    size_t smem_b;    // dynamic shared memory of pass B
    // workspace layout (bytes from the start of the tiled region)
    size_t o_win, o_counters, o_offs, o_code, o_wl, o_wr, o_partials, end;
};

static Geometry geometry(int64_t total, int B, int nb, int H, int W, int flavour, int sms) {
    Geometry g{};
    const int64_t HW = (int64_t)H * W;
    g.planes = nb * (flavour == CF_FLAVOUR_POL ? 2 : 1);
    if (B < 1 || B > (1 << 20) || HW >= (1ll << 28) || total < 0) return g;
    const int64_t pmax = (int64_t)(TILE_BYTES_MAX / (sizeof(float) * g.planes)) & ~3ll;
    if (pmax < 64) return g;
    const int64_t t0 = ceil_div(HW, pmax);
    // small batches: more (smaller) tiles so that every SM gets two items, but runs of >= 16 records per chunk and tile
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
    g.smem_b = 2 * (size_t)g.planes * P * sizeof(float);
    if (g.smem_b > 2 * TILE_BYTES_MAX) return g;
    const int64_t cap = sms;                         // one CTA per SM
    if (T > 2 * cap) return g;                       // a CTA may own at most two items of one window (kernel header)
    const int64_t items = (int64_t)B * T;
    g.grid_b = (int)(items < cap ? items : cap);
    int64_t mc = ceil_div(ceil_div(total > 0 ? total : 1, B), CHUNK);
    if (mc < 1) mc = 1;
    if (mc > 65536) mc = 65536;
    g.max_c = (int)mc;
    const int64_t a_items = (int64_t)B * mc;
    g.grid_a = (int)(a_items < 2 * (int64_t)sms ? a_items : 2 * (int64_t)sms);
    g.max_slots = (total > 0 ? total : 1) / CHUNK + B + 1;
    size_t o = 0;
    g.o_win = o;      o = align_up(o + (size_t)B * sizeof(WindowInfo), 256);
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

// ------------------------------------------------------------- window table ---
// begin / end / t0 / span of every window (event_process.py:39-44), and the arrival counters of pass B zeroed
__global__ void __launch_bounds__(256)
voxel_window_table_kernel(const double *__restrict__ ev, const int64_t *__restrict__ off, int B, WindowInfo *__restrict__ win,
                          int *__restrict__ counters) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    WindowInfo w;
    w.begin = __ldg(off + b);
    w.end = __ldg(off + b + 1);
    w.t0 = 0.0;
    w.span = 1.0;
    if (w.end > w.begin) {
        w.t0 = __ldg(ev + 4 * w.begin);
        w.span = __dsub_rn(__ldg(ev + 4 * (w.end - 1)), w.t0);
        if (w.span == 0.0) w.span = 1.0;  // event_process.py:43-44
    }
    win[b] = w;
    counters[b] = 0;
}

// ------------------------------------------------------------------ pass A ---
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ptx::smem_u32(dst)), "l"(src), "r"(bytes), "r"(ptx::smem_u32(bar)) : "memory");
}

// work items of pass A: (window b, chunk c) over a virtual [B][max_c] grid walked with stride gridDim.x; a window with
// more than max_c chunks (ragged batches: the host only knows the average) wraps around: c, c + max_c, c + 2 max_c, ...
struct PartItem {
    int64_t id;
    int b;
    int64_t c;
};
__device__ __forceinline__ bool part_first(PartItem &it, const WindowInfo *__restrict__ win, int B, int max_c, int64_t id0, int64_t stride) {
    for (int64_t id = id0; id < (int64_t)B * max_c; id += stride) {
        const int b = (int)(id / max_c);
        const int64_t c = id - (int64_t)b * max_c;
        if (c * CHUNK < win[b].end - win[b].begin) { it.id = id; it.b = b; it.c = c; return true; }
    }
    return false;
}
__device__ __forceinline__ bool part_next(PartItem &it, const WindowInfo *__restrict__ win, int B, int max_c, int64_t stride) {
    if ((it.c + max_c) * CHUNK < win[it.b].end - win[it.b].begin) { it.c += max_c; return true; }
    return part_first(it, win, B, max_c, it.id + stride, stride);
}

// Persistent: 2 CTAs per SM.  The raw 32-byte rows of a chunk arrive in shared memory by ONE bulk copy (cp.async.bulk,
// 64 KB), issued as soon as the previous chunk's rows have been decoded -- so the DRAM latency of chunk c+1 hides behind
// the counting sort and the record write-back of chunk c, and no thread holds event rows in registers across a load.
// (Round 2a loaded the rows with LDG: 2 CTAs x 4 rows per thread in flight per SM, a CTA lived 7.4 us per chunk and the
// pass ran at 2.3 TB/s.)
__global__ void __launch_bounds__(PART_THREADS, 2)
voxel_partition_kernel(const double *__restrict__ ev, const WindowInfo *__restrict__ win, int B, int max_c, int nb, int H, int W,
                       int flavour, int P, int T, uint32_t *__restrict__ offs,
                       uint32_t *__restrict__ rec_code, float *__restrict__ rec_wl, float *__restrict__ rec_wr) {
    __shared__ int s_hist[MAX_T + 1];
    __shared__ int s_warp[PART_THREADS / 32];
    __shared__ uint64_t s_full;
    extern __shared__ __align__(128) uint8_t s_dyn[];
    const double2 *s_raw = reinterpret_cast<const double2 *>(s_dyn);                 // [CHUNK][2] (t, x | y, p)
    uint32_t *s_code = reinterpret_cast<uint32_t *>(s_dyn + (size_t)CHUNK * 32);     // sorted records of the chunk
    float *s_wl = reinterpret_cast<float *>(s_code + CHUNK);
    float *s_wr = reinterpret_cast<float *>(s_code + 2 * CHUNK);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int planes_per_bin = flavour == CF_FLAVOUR_POL ? 2 : 1;

    if (tid == 0) {
        ptx::mbar_init(&s_full, 1);
        ptx::fence_barrier_init();
        CF_TRACE_AT(200);
    }
    PartItem cur;
    bool have = part_first(cur, win, B, max_c, blockIdx.x, gridDim.x);
    __syncthreads();
    if (have && tid == 0) {
        const WindowInfo wi = win[cur.b];
        const int64_t first = wi.begin + cur.c * CHUNK;
        const uint32_t bytes = (uint32_t)(min(wi.end, first + CHUNK) - first) * 32u;
        ptx::mbar_expect_tx(&s_full, bytes);
        bulk_load(s_dyn, ev + 4 * first, bytes, &s_full);
    }
    uint32_t phase = 0;
    while (have) {
        const WindowInfo wi = win[cur.b];
        Window w;
        w.b = cur.b; w.begin = wi.begin; w.end = wi.end; w.t0 = wi.t0; w.span = wi.span;
        const int64_t first = wi.begin + cur.c * CHUNK;
        const int count = (int)(min(wi.end, first + CHUNK) - first);
        const int64_t slot = wi.begin / CHUNK + cur.b + cur.c;
        PartItem nxt = cur;
        const bool have_next = part_next(nxt, win, B, max_c, gridDim.x);
        for (int t = tid; t <= T; t += PART_THREADS) s_hist[t] = 0;
        __syncthreads();
        ptx::mbar_wait_warp(&s_full, phase);
        phase ^= 1u;
        if (tid == 0) CF_TRACE_AT(201);
        uint32_t code[PER_THREAD], where[PER_THREAD];   // where = tile | rank << 16, 0xffffffff: dropped
        float wl[PER_THREAD], wr[PER_THREAD];
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k) {
            const int i = k * PART_THREADS + tid;
            where[k] = 0xffffffffu;
            if (i >= count) continue;
            const double2 lo2 = s_raw[2 * i], hi2 = s_raw[2 * i + 1];
            const Event e{lo2.x, lo2.y, hi2.x, hi2.y};
            const Binned bb = bin_event(e, w, nb, H, W, flavour);
            if (!bb.ok) continue;
            if (flavour == CF_FLAVOUR_TORCH) {
                weights_f32(bb, wl[k], wr[k]);
            } else {
                double dl, dr;
                weights_f64(bb, dl, dr);
                wl[k] = (float)dl;
                wr[k] = (float)dr;
            }
            const int pix = bb.y * W + bb.x;
            const int tile = pix / P;
            const int local = pix - tile * P;
            // cell index inside the tile's shared-memory image [planes][P]; bit 31: no right neighbour
            code[k] = (uint32_t)((bb.bin * planes_per_bin + bb.chan) * P + local) | (bb.bin + 1 < nb ? 0u : 0x80000000u);
            const int rank = atomicAdd(&s_hist[tile], 1);  // integer, shared memory: order-independent totals
            where[k] = (uint32_t)tile | ((uint32_t)rank << 16);
        }
        __syncthreads();   // the raw rows have been consumed
        if (tid == 0) {
            CF_TRACE_AT(202);
            if (have_next) {   // the next chunk's rows stream in behind the sort and the write-back of this one
                const WindowInfo wn = win[nxt.b];
                const int64_t nfirst = wn.begin + nxt.c * CHUNK;
                const uint32_t bytes = (uint32_t)(min(wn.end, nfirst + CHUNK) - nfirst) * 32u;
                ptx::mbar_expect_tx(&s_full, bytes);
                bulk_load(s_dyn, ev + 4 * nfirst, bytes, &s_full);
            }
        }
        // ---- exclusive scan of the T tile counts (in place; s_hist[T] = records of the chunk)
        {
            const int per = (T + PART_THREADS) / PART_THREADS;   // covers T + 1 entries
            const int s0 = tid * per, e0 = min(T + 1, s0 + per);
            int sum = 0;
            for (int i = s0; i < e0; ++i) sum += s_hist[i];
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
            for (int i = s0; i < e0; ++i) {
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
        uint32_t *o = offs + (size_t)slot * (T + 1);
        for (int t = tid; t <= T; t += PART_THREADS) o[t] = (uint32_t)s_hist[t];
        __syncthreads();   // the shared arrays are reused by the next chunk
        if (tid == 0) CF_TRACE_AT(204);
        cur = nxt;
        have = have_next;
    }
}

// ------------------------------------------------------------------ pass B ---
// fp32 add into the tile; feeds the telescoping statistics with (old, new)
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

__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(ACC_THREADS) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
}

// One CTA per SM, two warp groups, two tile buffers:
//   ACC (warps 0-15)   item k -> buffer k & 1: adds the tile's runs (CAS loop + telescoping statistics), publishes the
//                      tile's partial and bumps the window's arrival counter, hands the buffer to FIN, moves on to item
//                      k + 1 in the other buffer -- it never waits for a window, only for FIN to release a buffer;
//   FIN (warps 16-31)  waits for the buffer, then for the window's counter to reach T, derives mean / std from the T
//                      partials (fixed order), normalises out of shared memory, writes the grid once (128-bit streaming
//                      stores) and re-zeroes the buffer on the way.
// So the shared-memory atomics of item k + 1 overlap the HBM write of item k inside one SM, and the per-window barrier
// is waited for by warps that have nothing else to do.  Deadlock-free: a CTA's items are ordered by window and it owns
// at most two items of a window (T <= 2 x grid), so whoever an FIN group waits for is never blocked behind it.
template <int PRE>
__global__ void __launch_bounds__(2 * ACC_THREADS, 1)
voxel_tile_kernel(const uint32_t *__restrict__ rec_code, const float *__restrict__ rec_wl,
                  const float *__restrict__ rec_wr, const uint32_t *__restrict__ offs,
                  const WindowInfo *__restrict__ win, int *counters, Partial *partials, int B, int planes,
                  int right_stride /* cells between a bin and the next one inside the tile */, int64_t HW, int P,
                  int T, float hot_thr, float *__restrict__ out) {
    extern __shared__ __align__(16) float tiles[];   // [2][planes][P]
    __shared__ Partial s_part[ACC_WARPS];
    __shared__ float s_a, s_b;
    __shared__ int s_identity;
    __shared__ uint64_t acc_done[2], fin_done[2];
    const int tid = threadIdx.x;
    const int cells = planes * P;
    const int64_t items = (int64_t)B * T;

    for (int i = tid; i < 2 * cells / 4; i += 2 * ACC_THREADS) reinterpret_cast<float4 *>(tiles)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        for (int k = 0; k < 2; ++k) { ptx::mbar_init(&acc_done[k], 1); ptx::mbar_init(&fin_done[k], 1); }
        ptx::fence_barrier_init();
    }
    __syncthreads();

    if (tid < ACC_THREADS) {
        // ------------------------------------------------------------------ ACC group
        const int lane = tid & 31, warp = tid >> 5;
        int k = 0;
        for (int64_t item = blockIdx.x; item < items; item += gridDim.x, ++k) {
            const int buf = k & 1;
            float *tile = tiles + (size_t)buf * cells;
            const int b = (int)(item / T), t = (int)(item - (int64_t)b * T);
            const WindowInfo wi = win[b];
            const int nchunks = (int)((wi.end - wi.begin + CHUNK - 1) / CHUNK);
            const int64_t slot0 = wi.begin / CHUNK + b;
            if (k >= 2) ptx::mbar_wait_warp(&fin_done[buf], (uint32_t)((k >> 1) - 1) & 1u);   // FIN wrote item k-2 out, buffer is zero again
            if (tid == 0) CF_TRACE_AT(8 * k + 0);
            // warp <-> 4 chunks at a time: the run bounds of 4 chunks, then two records per lane and chunk, in flight together
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
                uint32_t code[2 * U];
                float wa[2 * U], wb[2 * U];
                bool ok[2 * U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int q = 2 * u + h;
                        ok[q] = lo[u] + lane + 32 * h < hi[u];
                        const size_t at = (size_t)(wi.begin + (int64_t)(c0 + u) * CHUNK) + lo[u] + lane + 32 * h;
                        if (ok[q]) { code[q] = __ldg(rec_code + at); wa[q] = __ldg(rec_wl + at); wb[q] = __ldg(rec_wr + at); }
                    }
                }
#pragma unroll
                for (int q = 0; q < 2 * U; ++q) {
                    if (ok[q]) {
                        const int cell = (int)(code[q] & 0x7fffffffu);
                        tile_add<PRE == CF_PRE_STD>(tile + cell, wa[q], hot_thr, d_sum, d_sq, d_nnz);
                        if (!(code[q] >> 31)) tile_add<PRE == CF_PRE_STD>(tile + cell + right_stride, wb[q], hot_thr, d_sum, d_sq, d_nnz);
                    }
                }
#pragma unroll 1
                for (int u = 0; u < U; ++u) {  // runs longer than two warps (hot tiles): the rest, 32 at a time
                    if (c0 + u >= nchunks) break;
                    const size_t base = (size_t)(wi.begin + (int64_t)(c0 + u) * CHUNK);
                    for (int i = lo[u] + 64 + lane; i < hi[u]; i += 32) {
                        const uint32_t cd = __ldg(rec_code + base + i);
                        const float x = __ldg(rec_wl + base + i), y = __ldg(rec_wr + base + i);
                        const int cell = (int)(cd & 0x7fffffffu);
                        tile_add<PRE == CF_PRE_STD>(tile + cell, x, hot_thr, d_sum, d_sq, d_nnz);
                        if (!(cd >> 31)) tile_add<PRE == CF_PRE_STD>(tile + cell + right_stride, y, hot_thr, d_sum, d_sq, d_nnz);
                    }
                }
            }
            if (PRE != CF_PRE_NONE) {
                // ---- this tile's share of the window statistics (hot pixels filtered first, event_process.py:196-198)
                double sum = 0.0, sumsq = 0.0;
                long long nnz = 0;
                float mn = INFINITY, mx = -INFINITY;
                if (PRE == CF_PRE_STD) {          // telescoped out of the scatter
                    sum = d_sum; sumsq = d_sq; nnz = d_nnz;
                } else {                          // min / max do not telescope: one pass over the tile's valid cells
                    group_sync(1);                // every add of the group has landed
                    const int valid = (int)min((int64_t)P, HW - (int64_t)t * P);
                    for (int pl = 0; pl < planes; ++pl) {
                        const float *row = tile + pl * P;
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
            }
            group_sync(1);   // all adds (and partials) of this item are in shared memory
            if (tid == 0) {
                if (PRE != CF_PRE_NONE) {
                    Partial p = s_part[0];
                    for (int q = 1; q < ACC_WARPS; ++q) {
                        p.sum += s_part[q].sum; p.sumsq += s_part[q].sumsq; p.nnz += s_part[q].nnz;
                        p.mn = fminf(p.mn, s_part[q].mn); p.mx = fmaxf(p.mx, s_part[q].mx);
                    }
                    partials[(size_t)b * T + t] = p;
                    __threadfence();
                    atomicAdd(counters + b, 1);
                }
                CF_TRACE_AT(8 * k + 1);
                mbar_arrive(&acc_done[buf]);
            }
            group_sync(1);   // s_part is reused by the next item
        }
    } else {
        // ------------------------------------------------------------------ FIN group
        const int ftid = tid - ACC_THREADS, lane = ftid & 31, warp = ftid >> 5;
        int k = 0;
        for (int64_t item = blockIdx.x; item < items; item += gridDim.x, ++k) {
            const int buf = k & 1;
            float *tile = tiles + (size_t)buf * cells;
            const int b = (int)(item / T), t = (int)(item - (int64_t)b * T);
            const int px0 = t * P;
            const int valid = (int)min((int64_t)P, HW - px0);   // pixels of this tile inside the grid (> 0 by construction)
            ptx::mbar_wait_warp(&acc_done[buf], (uint32_t)(k >> 1) & 1u);
            if (ftid == 0) CF_TRACE_AT(8 * k + 2);
            float a = 0.f, inv = 1.f;
            bool identity = true;
            if (PRE != CF_PRE_NONE) {
                if (ftid == 0) {
                    // every tile of window b is published by an ACC group that never waits on a window: bounded spin
                    unsigned spins = 0;
                    while (*reinterpret_cast<volatile int *>(counters + b) < T) {
                        __nanosleep(32);
                        if (++spins > (1u << 24)) __trap();
                    }
                    __threadfence();
                }
                group_sync(2);
                if (warp == 0) {  // every CTA of the window combines the T partials in the same fixed order
                    const Partial *p = partials + (size_t)b * T;
                    double ts = 0.0, tq = 0.0;
                    long long tn = 0;
                    float tmn = INFINITY, tmx = -INFINITY;
                    for (int q = lane; q < T; q += 32) {
                        const volatile Partial *pp = p + q;   // written by other CTAs of this launch: bypass L1
                        ts += pp->sum; tq += pp->sumsq; tn += pp->nnz;
                        tmn = fminf(tmn, pp->mn); tmx = fmaxf(tmx, pp->mx);
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
                group_sync(2);
                a = s_a; inv = s_b; identity = s_identity != 0;
            }
            if (ftid == 0) CF_TRACE_AT(8 * k + 3);
            // ---- normalise out of shared memory (the fp32 map of voxel_normalise_kernel), write the final grid once,
            //      leave the buffer zeroed for item k + 2
            auto norm = [&](float raw) -> float {
                if (PRE == CF_PRE_NONE) return raw;
                const float v = hot_filter(raw, hot_thr);
                if (identity) return v;
                const float r = (v - a) * inv;
                return (PRE == CF_PRE_STD && v == 0.f) ? 0.f : r;
            };
            float *ob = out + (size_t)b * planes * HW + px0;
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((HW & 3) == 0) {   // plane starts and tile starts are 16-byte aligned
                const int v4 = valid / 4;   // valid % 4 == 0: HW % 4 == 0 and P % 4 == 0
                for (int pl = 0; pl < planes; ++pl) {
                    float4 *src = reinterpret_cast<float4 *>(tile + pl * P);
                    float4 *dst = reinterpret_cast<float4 *>(ob + (size_t)pl * HW);
                    for (int j = ftid; j < v4; j += ACC_THREADS) {
                        const float4 q = src[j];
                        src[j] = zero4;
                        st_cs4(dst + j, make_float4(norm(q.x), norm(q.y), norm(q.z), norm(q.w)));
                    }
                }
            } else {
                for (int pl = 0; pl < planes; ++pl)
                    for (int j = ftid; j < valid; j += ACC_THREADS) {
                        const float v = tile[pl * P + j];
                        tile[pl * P + j] = 0.f;
                        st_cs(ob + (size_t)pl * HW + j, norm(v));
                    }
            }
            group_sync(2);   // the whole buffer has been read and re-zeroed; s_a / s_b may be rewritten
            if (ftid == 0) {
                CF_TRACE_AT(8 * k + 4);
                mbar_arrive(&fin_done[buf]);
            }
        }
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
    WindowInfo *win = reinterpret_cast<WindowInfo *>(w8 + g.o_win);
    int *counters = reinterpret_cast<int *>(w8 + g.o_counters);
    uint32_t *offs = reinterpret_cast<uint32_t *>(w8 + g.o_offs);
    uint32_t *rec_code = reinterpret_cast<uint32_t *>(w8 + g.o_code);
    float *rec_wl = reinterpret_cast<float *>(w8 + g.o_wl);
    float *rec_wr = reinterpret_cast<float *>(w8 + g.o_wr);
    Partial *partials = reinterpret_cast<Partial *>(w8 + g.o_partials);

    const void *fn = preprocess == CF_PRE_STD ? reinterpret_cast<const void *>(voxel_tile_kernel<CF_PRE_STD>)
                     : preprocess == CF_PRE_MAXMIN ? reinterpret_cast<const void *>(voxel_tile_kernel<CF_PRE_MAXMIN>)
                                                   : reinterpret_cast<const void *>(voxel_tile_kernel<CF_PRE_NONE>);
    int dev = 0;
    CF_CUDA(cudaGetDevice(&dev));
    static bool opt_in[64] = {};
    if (!opt_in[dev & 63]) {
        CF_CUDA(cudaFuncSetAttribute(voxel_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PART_SMEM));
        CF_CUDA(cudaFuncSetAttribute(voxel_tile_kernel<CF_PRE_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * TILE_BYTES_MAX)));
        CF_CUDA(cudaFuncSetAttribute(voxel_tile_kernel<CF_PRE_STD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * TILE_BYTES_MAX)));
        CF_CUDA(cudaFuncSetAttribute(voxel_tile_kernel<CF_PRE_MAXMIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * TILE_BYTES_MAX)));
        opt_in[dev & 63] = true;
    }
    int resident = 0;
    CF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, fn, 2 * ACC_THREADS, g.smem_b));
    const int64_t cap = (int64_t)resident * sm_count();
    if (2 * cap < g.T || cap < 1) return 1;   // cannot keep a window's tiles within two items per CTA: the L2-atomic path takes over

    voxel_window_table_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, stream>>>(events, offsets, B, win, counters);
    CF_LAUNCH_CHECK("voxel_window_table_kernel");
    voxel_partition_kernel<<<(unsigned)g.grid_a, PART_THREADS, PART_SMEM, stream>>>(
        events, win, B, g.max_c, nb, H, W, flavour, g.P, g.T, offs, rec_code, rec_wl, rec_wr);
    CF_LAUNCH_CHECK("voxel_partition_kernel");

    const int planes = g.planes;
    const int right_stride = (flavour == CF_FLAVOUR_POL ? 2 : 1) * g.P;
    const int64_t HW = (int64_t)H * W;
    int P = g.P, T = g.T;
    const uint32_t *c_code = rec_code;
    const float *c_wl = rec_wl, *c_wr = rec_wr;
    const uint32_t *c_offs = offs;
    const WindowInfo *c_win = win;
    int Bv = B, rs = right_stride, pl = planes;
    int64_t hw = HW;
    float thr = hot_thr;
    void *args[] = {&c_code, &c_wl, &c_wr, &c_offs, &c_win, &counters, &partials, &Bv, &pl, &rs, &hw, &P, &T, &thr, &out};
    const int grid_b = (int)(g.grid_b < cap ? g.grid_b : cap);
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid_b), dim3(2 * ACC_THREADS), args, g.smem_b, stream);
    count_launch("voxel_tile_kernel");
    if (e != cudaSuccess) {
        set_error("cooperative launch of voxel_tile_kernel (%d CTAs, %zu B smem) failed: %s", grid_b, g.smem_b,
                  cudaGetErrorString(e));
        return CF_ERR_CUDA;
    }
    return CF_OK;
}

}  // namespace cf
