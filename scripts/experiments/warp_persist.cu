// Persistent, warp-specialised flow-guided warp for multi-channel tensors (the CISTA-LSTC sparse codes).
//
// Same maths as warp.cu (utils/flow_utils.py:83-120,153-190 of the reference); different data movement.
// What the measurements on the B200 said (scripts/scale_bench.py, timelines from scripts/warp_trace.py,
// ncu captures under profiles/):
//   * direct gather (warp.cu): every in-flight byte sits in a register of a stalled thread, 4 mostly
//     redundant tap loads per output                                         -> ~30 % of HBM bandwidth;
//   * one CTA per (32x16 tile, 32 channels) staging a fixed 48x24 tap box through TMA (warp_tma.cu):
//     43-57 %.  At 480x640 it moves 5.9 TB/s between L2 and the SMs -- the same ceiling a plain copy
//     reaches (6.5 TB/s read+write): the kernel is bound by L2<->SM traffic, and the 48x24 box of a
//     32x16 tile fetches 2.25x the tile.  The ceiling of that design is 2/(1+2.25) = 61 %.
//   * so the lever is the over-fetch, not latency: this kernel uses 64x32 pixel tiles (halo of a
//     smooth flow: a few pixels -> box/tile = 1.15-1.4) and picks, per tile, the smallest box of a 3x3
//     menu of TMA tensor maps (widths 68/72/80 x heights 34/36/40) that holds all taps of the tile.
// Structure: one CTA per SM, alive for the whole launch.  The work is a flat stream of CHUNKS (batch
// item, 64x32 tile, 2 channels), split evenly and contiguously over the CTAs.
//   - 4 producer warps: reduce the bounding box of the taps of a tile one tile AHEAD of the chunks
//     being issued; one thread then issues one cp.async.bulk.tensor.4d per chunk into a shared-memory
//     ring (8 stages x 25 KB), completing on the stage's mbarrier;
//   - 16 consumer warps: thread <-> 4 pixels of the tile (2 rows x 2 half-rows); sample positions once
//     per tile; per chunk all 32 shared-memory tap loads, then 8 x (4 FMA + 1 coalesced streaming
//     store); release the stage.  One lane per warp polls the barrier (32 spinning lanes compete with
//     the working warps for the shared-memory pipe);
//   - tiles whose taps do not fit the largest box (flow discontinuities) are gathered directly from
//     global memory by the consumers, chunk by chunk, in the same ring order;
//   - 2 image warps: the 1-channel full-resolution image of the per-frame step (direct gather),
//     concurrently with the codes pipeline.
// Shared memory is declared aligned instead of rounding the base at run time: that keeps the pointers in
// the shared address space (LDS; the rounded pointer of round-1a compiled to generic LD.E).
// Needs a 16-byte aligned row pitch (W % 4 == 0) for the tensor maps; other shapes use warp.cu.
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include "tma.cuh"
#include "warp_common.cuh"

namespace cf {

namespace wp {
constexpr int TW = 64, TH = 16;          // output tile
constexpr int CC = 4;                    // channels per chunk
constexpr int NBW = 3, NBH = 3;          // box menu
__host__ __device__ constexpr int box_w(int i) { return i == 0 ? 68 : (i == 1 ? 72 : 80); }
__host__ __device__ constexpr int box_h(int i) { return i == 0 ? 18 : (i == 1 ? 20 : 24); }
constexpr int MAX_BW = 80, MAX_BH = 24;
constexpr int CONSUMER_WARPS = 16, PRODUCER_WARPS = 4, IMAGE_WARPS = 2;
constexpr int PRODUCER_THREADS = 32 * PRODUCER_WARPS, IMAGE_THREADS = 32 * IMAGE_WARPS;
constexpr int THREADS = 32 * (CONSUMER_WARPS + PRODUCER_WARPS + IMAGE_WARPS);  // 704
constexpr int MODE_SMEM = 0, MODE_DIRECT = 1;
constexpr int STAGE_FLOATS = CC * MAX_BH * MAX_BW;   // 7680
constexpr int STAGE_BYTES = STAGE_FLOATS * 4;        // 30 720
constexpr int STAGES = 7;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 512 /*barriers + meta + bbox exchange*/;

struct Meta { int bx, by, mode, bw, chan_stride, pad0, pad1, pad2; };   // 32 bytes
struct Box { int bx, by, wi, hi, fits; };

struct alignas(64) Maps { CUtensorMap m[NBH][NBW]; };

struct Params {
    WarpJob ji; int with_image;
    WarpJob jz; int tiles_x, tiles;     // tiles per batch item
    int cpt;                            // chunks per tile = ceil(C / CC)
    int B;
    int total_chunks;
    const float *flow; int fH, fW; float sign;
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void producer_sync() {  // named barrier 1: the 4 producer warps only
    asm volatile("bar.sync 1, %0;" ::"n"(PRODUCER_THREADS) : "memory");
}

#ifdef CF_TRACE
// experiment builds only (scripts/warp_trace.py): per-CTA timeline in globaltimer ns
__device__ unsigned long long *g_trace = nullptr;
constexpr int TRACE_SLOTS = 256;
__device__ __forceinline__ void trace(int slot) {
    if (g_trace && slot < TRACE_SLOTS) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_trace[(size_t)blockIdx.x * TRACE_SLOTS + slot] = t;
    }
}
#define CF_TRACE_AT(slot) wp::trace(slot)
#else
#define CF_TRACE_AT(slot) ((void)0)
#endif

struct TapPos { int x0, y0, x1, y1; };
__device__ __forceinline__ TapPos tap_pos(const Taps &t, int W) {
    TapPos p;
    p.y0 = t.o00 / W; p.x0 = t.o00 - p.y0 * W;
    p.x1 = t.o01 - p.y0 * W;  // clamped to W-1
    p.y1 = t.o10 / W;         // clamped to H-1
    return p;
}

// Bounding box of all taps of tile `tg` (global tile index) and the smallest menu box that holds it.
// 128 producer threads: thread <-> column pair (lane, lane + 32) x 8 rows (warp); partial results meet in
// shared memory.  Kept out of line: one copy in the instruction cache, off the issue loop.
__device__ __noinline__ Box tile_box(const Params &P, int tg, int pw, int lane, int (*xch)[4]) {
    const WarpJob &jz = P.jz;
    const int H = jz.H, W = jz.W;
    const int b = tg / P.tiles;
    const int tile = tg - b * P.tiles;
    const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
    const float *fb = P.flow + (size_t)b * 2 * P.fH * P.fW;
    int mnx = INT_MAX, mny = INT_MAX, mxx = -1, mxy = -1;
    constexpr int RPW = TH / PRODUCER_WARPS;  // 4 rows per warp
#pragma unroll 2
    for (int i = 0; i < 2 * RPW; ++i) {
        // clamped duplicates do not change the box
        const int x = min(tx * TW + lane + 32 * (i & 1), W - 1);
        const int y = min(ty * TH + pw * RPW + (i >> 1), H - 1);
        const float2 uv = flow_at(fb, x, y, W, P.fH, P.fW, jz.half != 0, jz.sy, jz.sx);
        const TapPos tp = tap_pos(make_taps(uv.x, uv.y, x, y, H, W, P.sign), W);
        mnx = min(mnx, tp.x0); mny = min(mny, tp.y0);
        mxx = max(mxx, tp.x1); mxy = max(mxy, tp.y1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    producer_sync();  // previous exchange fully read
    if (lane == 0) { xch[pw][0] = mnx; xch[pw][1] = mny; xch[pw][2] = mxx; xch[pw][3] = mxy; }
    producer_sync();
#pragma unroll
    for (int k = 0; k < PRODUCER_WARPS; ++k) {
        mnx = min(mnx, xch[k][0]); mny = min(mny, xch[k][1]);
        mxx = max(mxx, xch[k][2]); mxy = max(mxy, xch[k][3]);
    }
    Box bb;
    bb.bx = mnx & ~3;   // 16-byte aligned origin keeps the TMA requests sector-aligned
    bb.by = mny;
    const int need_w = mxx - bb.bx + 1, need_h = mxy - bb.by + 1;
    bb.wi = need_w <= box_w(0) ? 0 : (need_w <= box_w(1) ? 1 : 2);
    bb.hi = need_h <= box_h(0) ? 0 : (need_h <= box_h(1) ? 1 : 2);
    bb.fits = need_w <= MAX_BW && need_h <= MAX_BH;
    return bb;
}

// Direct gather of one chunk of one pixel (tiles that do not fit any box); out of line, rare.
__device__ __noinline__ void direct_pixel(const float *__restrict__ img_c, float *__restrict__ out_c, Taps t, int p,
                                          int nch, size_t plane) {
    for (int c = 0; c < nch; ++c) {
        const float *s = img_c + (size_t)c * plane;
        float r = __ldg(s + t.o00) * t.w00;
        r += __ldg(s + t.o01) * t.w01;
        r += __ldg(s + t.o10) * t.w10;
        r += __ldg(s + t.o11) * t.w11;
        st_cs(out_c + (size_t)c * plane + p, r);
    }
}
}  // namespace wp

__global__ void __launch_bounds__(wp::THREADS, 1)
warp_persist_kernel(const __grid_constant__ wp::Maps maps, const __grid_constant__ wp::Params P) {
    using namespace wp;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + STAGES * STAGE_FLOATS);
    uint64_t *empty = full + STAGES;
    Meta *meta = reinterpret_cast<Meta *>(empty + STAGES);
    int (*xch)[4] = reinterpret_cast<int (*)[4]>(meta + STAGES);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WarpJob &jz = P.jz;
    const int H = jz.H, W = jz.W, C = jz.C;
    const size_t plane = (size_t)H * W;

    if (tid == 0) {
        CF_TRACE_AT(0);
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], CONSUMER_WARPS);
        }
        ptx::fence_barrier_init();
    }
    __syncthreads();

    const int k0 = (int)((long long)blockIdx.x * P.total_chunks / gridDim.x);
    const int k1 = (int)((long long)(blockIdx.x + 1) * P.total_chunks / gridDim.x);

    if (warp >= CONSUMER_WARPS + PRODUCER_WARPS) {
        // ================================ image warps =====================================
        if (!P.with_image) return;
        const WarpJob &ji = P.ji;
        const int Hi = ji.H, Wi = ji.W;
        const long long hw = (long long)Hi * Wi, total = hw * P.B;
        // CTA ranges in units of 32 pixels keep the warps' accesses on 128-byte lines
        const long long units = (total + 31) / 32;
        const long long p0 = (long long)blockIdx.x * units / gridDim.x * 32;
        const long long p1 = min(total, (long long)(blockIdx.x + 1) * units / gridDim.x * 32);
        const int t = tid - 32 * (CONSUMER_WARPS + PRODUCER_WARPS);
        constexpr int U = 2;
        for (long long base = p0 + t; base < p1; base += (long long)U * IMAGE_THREADS) {
            int bb[U], pp[U], xx[U], yy[U];
            float fu[U], fv[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long idx = base + (long long)u * IMAGE_THREADS;
                ok[u] = idx < p1;
                const long long id = ok[u] ? idx : p0;
                bb[u] = (int)(id / hw);
                pp[u] = (int)(id - (long long)bb[u] * hw);
                yy[u] = pp[u] / Wi;
                xx[u] = pp[u] - yy[u] * Wi;
                const float *fb = P.flow + (size_t)bb[u] * 2 * P.fH * P.fW;  // image resolution == flow resolution
                fu[u] = __ldg(fb + pp[u]);
                fv[u] = __ldg(fb + hw + pp[u]);
            }
            Taps tp[U];
#pragma unroll
            for (int u = 0; u < U; ++u) tp[u] = make_taps(fu[u], fv[u], xx[u], yy[u], Hi, Wi, P.sign);
            for (int c = 0; c < ji.C; ++c) {
                float v[U][4];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float *s = ji.img + ((size_t)bb[u] * ji.C + c) * hw;
                    v[u][0] = __ldg(s + tp[u].o00); v[u][1] = __ldg(s + tp[u].o01);
                    v[u][2] = __ldg(s + tp[u].o10); v[u][3] = __ldg(s + tp[u].o11);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    float r = v[u][0] * tp[u].w00;
                    r += v[u][1] * tp[u].w01;
                    r += v[u][2] * tp[u].w10;
                    r += v[u][3] * tp[u].w11;
                    if (ok[u]) st_cs(ji.out + ((size_t)bb[u] * ji.C + c) * hw + pp[u], r);
                }
            }
        }
        if (t == 0) CF_TRACE_AT(1);
        return;
    }

    if (warp >= CONSUMER_WARPS) {
        // ================================ producers =======================================
        if (k0 >= k1) return;
        const int pw = warp - CONSUMER_WARPS, pt = tid - 32 * CONSUMER_WARPS;
        int tg = k0 / P.cpt, ck = k0 - tg * P.cpt;  // tile (global index) and chunk inside the tile
        const int last_tg = (k1 - 1) / P.cpt;
        Box cur = tile_box(P, tg, pw, lane, xch), nxt = cur;
        bool have_next = false;
        int b = tg / P.tiles;
        int s = 0;
        uint32_t ph = 0;
        for (int n = 0; n < k1 - k0; ++n) {
            if (pt == 0) {  // only the issuing thread needs the slot
                ptx::mbar_wait(&empty[s], ph ^ 1u);
                CF_TRACE_AT(8 + 4 * n);
                const int bw = box_w(cur.wi), bh = box_h(cur.hi);
                meta[s] = Meta{cur.bx, cur.by, cur.fits ? MODE_SMEM : MODE_DIRECT, bw, bw * bh, 0, 0, 0};
                if (cur.fits) {
                    ptx::mbar_expect_tx(&full[s], (uint32_t)(CC * bw * bh * 4));
                    ptx::tma_load_4d(ring + s * STAGE_FLOATS, &maps.m[cur.hi][cur.wi], cur.bx, cur.by, ck * CC, b, &full[s]);
                } else {
                    mbar_arrive(&full[s]);
                }
            }
            // one tile of look-ahead: the next tile's box is reduced while this tile's chunks are in flight
            if (!have_next && tg < last_tg) {
                nxt = tile_box(P, tg + 1, pw, lane, xch);
                have_next = true;
            }
            if (++s == STAGES) { s = 0; ph ^= 1u; }
            if (++ck == P.cpt) {
                ck = 0;
                ++tg;
                b = tg / P.tiles;
                if (n + 1 < k1 - k0) cur = have_next ? nxt : tile_box(P, tg, pw, lane, xch);
                have_next = false;
            }
        }
        return;
    }

    // ==================================== consumers =======================================
    if (k0 >= k1) return;
    int tg = k0 / P.cpt, ck = k0 - tg * P.cpt;
    bool new_tile = true;
    // thread <-> pixels (row warp, column lane + 32*j) of the tile, j = 0..1
    float tw[2][4];   // bilinear weights
    int xy[2];        // top-left tap: x0 | y0 << 16
    int pixf[2];      // output offset inside a channel plane | (x1 - x0) << 29 | (y1 - y0) << 30; < 0: outside the image
    int b = 0;
    int s = 0;
    uint32_t ph = 0;
    for (int n = 0; n < k1 - k0; ++n) {
        if (new_tile) {
            new_tile = false;
            b = tg / P.tiles;
            const int tile = tg - b * P.tiles;
            const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
            const float *fb = P.flow + (size_t)b * 2 * P.fH * P.fW;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int x = tx * TW + lane + 32 * j, y = ty * TH + warp;
                const int xc = min(x, W - 1), yc = min(y, H - 1);   // dead pixels shadow a live one, store nothing
                const float2 uv = flow_at(fb, xc, yc, W, P.fH, P.fW, jz.half != 0, jz.sy, jz.sx);
                const Taps t = make_taps(uv.x, uv.y, xc, yc, H, W, P.sign);
                const TapPos tp = tap_pos(t, W);
                tw[j][0] = t.w00; tw[j][1] = t.w01; tw[j][2] = t.w10; tw[j][3] = t.w11;
                xy[j] = tp.x0 | (tp.y0 << 16);
                pixf[j] = (x < W && y < H) ? ((yc * W + xc) | ((tp.x1 - tp.x0) << 29) | ((tp.y1 - tp.y0) << 30)) : -1;
            }
        }
        if (tid == 0) CF_TRACE_AT(8 + 4 * n + 1);
        ptx::mbar_wait_warp(&full[s], ph);
        if (tid == 0) CF_TRACE_AT(8 + 4 * n + 2);
        const Meta m = meta[s];
        const int c0 = ck * CC;
        const int nch = min(CC, C - c0);
        float *out_c = jz.out + ((size_t)b * C + c0) * plane;
        if (m.mode == MODE_SMEM) {
            const float *st = ring + s * STAGE_FLOATS;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float v[CC][4];  // the 16 tap loads of a pixel before its first store
                const int dx = (pixf[j] >> 29) & 1, dyo = ((pixf[j] >> 30) & 1) * m.bw;
                const float *s0 = st + (((xy[j] >> 16) - m.by) * m.bw + (xy[j] & 0xffff) - m.bx);
                const float *s1 = s0 + dx, *s2 = s0 + dyo, *s3 = s2 + dx;
#pragma unroll
                for (int c = 0; c < CC; ++c) {
                    v[c][0] = s0[c * m.chan_stride];
                    v[c][1] = s1[c * m.chan_stride];
                    v[c][2] = s2[c * m.chan_stride];
                    v[c][3] = s3[c * m.chan_stride];
                }
                if (pixf[j] >= 0) {
                    float *o = out_c + (pixf[j] & 0x1fffffff);
#pragma unroll
                    for (int c = 0; c < CC; ++c) {
                        // ATen order: nw, ne, sw, se accumulated left to right
                        float a = v[c][0] * tw[j][0];
                        a += v[c][1] * tw[j][1];
                        a += v[c][2] * tw[j][2];
                        a += v[c][3] * tw[j][3];
                        if (c < nch) st_cs(o, a);
                        o += plane;
                    }
                }
            }
        } else {
            const float *img_c = jz.img + ((size_t)b * C + c0) * plane;
#pragma unroll 1
            for (int j = 0; j < 2; ++j) {
                if (pixf[j] >= 0) {
                    const int dx = (pixf[j] >> 29) & 1, dyo = ((pixf[j] >> 30) & 1) * W;
                    Taps t;
                    t.o00 = (xy[j] >> 16) * W + (xy[j] & 0xffff); t.o01 = t.o00 + dx;
                    t.o10 = t.o00 + dyo; t.o11 = t.o10 + dx;
                    t.w00 = tw[j][0]; t.w01 = tw[j][1]; t.w10 = tw[j][2]; t.w11 = tw[j][3];
                    direct_pixel(img_c, out_c, t, pixf[j] & 0x1fffffff, nch, plane);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (tid == 0) CF_TRACE_AT(8 + 4 * n + 3);
        if (++s == STAGES) { s = 0; ph ^= 1u; }
        if (++ck == P.cpt) { ck = 0; ++tg; new_tile = true; }
    }
    if (tid == 0) CF_TRACE_AT(2);
}

#ifdef CF_TRACE
extern "C" __attribute__((visibility("default"))) int cf_trace_buffer(unsigned long long *buf) {
    return (int)cudaMemcpyToSymbol(wp::g_trace, &buf, sizeof(buf));
}
extern "C" __attribute__((visibility("default"))) int cf_trace_slots(void) { return wp::TRACE_SLOTS; }
#endif

// CF_OK / error, or 1 when the staged path does not apply (caller falls back to warp_tma.cu / warp.cu)
int launch_warp_persist(const WarpJob &ji, bool with_image, const WarpJob &jz, const float *flow, int fH, int fW,
                        float sign, int B, cudaStream_t stream) {
    using namespace wp;
    static const char *env = getenv("CF_WARP_PATH");  // experiments: "direct" | "tma" (round-1a kernel) | default
    if (env && (!strcmp(env, "direct") || !strcmp(env, "tma"))) return 1;
    if (jz.C < CC || jz.W % 4 != 0 || !aligned16(jz.img)) return 1;
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return 1;
    Params P{};
    P.ji = ji;
    P.with_image = with_image ? 1 : 0;
    P.jz = jz;
    P.tiles_x = (int)ceil_div(jz.W, TW);
    P.tiles = P.tiles_x * (int)ceil_div(jz.H, TH);
    P.cpt = (int)ceil_div(jz.C, CC);
    P.B = B;
    const long long chunks = (long long)B * P.tiles * P.cpt;
    if (chunks >= (1ll << 31) || (long long)jz.H * jz.W >= (1ll << 29) || jz.W > 65535 || jz.H > 32767) return 1;
    P.total_chunks = (int)chunks;
    P.flow = flow; P.fH = fH; P.fW = fW; P.sign = sign;
    Maps maps;
    cuuint64_t dims[4] = {(cuuint64_t)jz.W, (cuuint64_t)jz.H, (cuuint64_t)jz.C, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)jz.W * 4, (cuuint64_t)jz.W * jz.H * 4, (cuuint64_t)jz.W * jz.H * jz.C * 4};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int hi = 0; hi < NBH; ++hi)
        for (int wi = 0; wi < NBW; ++wi) {
            cuuint32_t box[4] = {(cuuint32_t)box_w(wi), (cuuint32_t)box_h(hi), CC, 1};
            CUresult r = enc(&maps.m[hi][wi], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(jz.img), dims, strides,
                             box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            CF_REQUIRE(r == CUDA_SUCCESS, CF_ERR_CUDA, "cuTensorMapEncodeTiled (warp) failed with CUresult %d", (int)r);
        }
    int dev = 0;
    CF_CUDA(cudaGetDevice(&dev));
    static bool opt_in[64] = {};
    if (!opt_in[dev & 63]) {
        CF_CUDA(cudaFuncSetAttribute(warp_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        opt_in[dev & 63] = true;
    }
    long long grid = sm_count();
    if (grid > P.total_chunks) grid = P.total_chunks;
    if (grid < 1) grid = 1;
    warp_persist_kernel<<<(unsigned)grid, THREADS, SMEM_BYTES, stream>>>(maps, P);
    CF_LAUNCH_CHECK("warp_persist_kernel");
    return CF_OK;
}

}  // namespace cf
