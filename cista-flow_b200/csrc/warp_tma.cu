// TMA-staged flow-guided warp for multi-channel tensors (the CISTA-LSTC sparse codes).
//
// Same maths as warp.cu (utils/flow_utils.py:83-120,153-190 of the reference); different data
// movement.  The direct gather keeps every in-flight byte in a register of a stalled thread and
// issues 4 (mostly redundant) tap loads per output, which capped it at ~30 % of HBM bandwidth.
// Here one CTA owns a 32x16 pixel tile:
//   1. every thread computes the sample position of its 2 pixels (flow x0.5 down-sampling fused)
//      and the CTA reduces the bounding box of all taps of the tile;
//   2. if the box fits 48x24 source pixels (it does wherever the flow is smooth), an elected thread
//      streams the box of 8 channels at a time with cp.async.bulk.tensor.4d (TMA, zero-filled
//      outside the image) into a 2-stage shared-memory ring (3 CTAs per SM) -- bytes in flight are
//      now bounded by shared memory (222 KB per SM), not by registers;
//   3. the threads gather their 4 taps from shared memory (conflict-free: consecutive lanes read
//      consecutive columns) and write 128-byte rows with streaming stores.
//   Tiles whose box does not fit (flow discontinuities) fall back to the direct gather in place.
// Variants measured on the B200 and NOT adopted (scripts/scale_bench.py; round-1 notes in DESIGN.md):
//   * persistent warp-specialised kernel, one CTA per SM, producer warps + mbarrier ring + consumer
//     warps (scripts/experiments/warp_persist.cu): 18-30 % -- with 8-16 consumer warps per SM the
//     per-chunk issue latency (2-4 warps per scheduler, i-cache misses of three code roles, and, under
//     a tight register cap, local-memory spills that miss the ~10 KB of L1 left beside a 215 KB ring)
//     bounds it, not memory (timelines: scripts/warp_trace.py);
//   * per-row cp.async.bulk (UBLKCP) for shapes without a 16-byte row pitch: ~70 cycles per issued
//     copy, serialised per lane;
//   * a second, smaller source box (40x20, 1.56x instead of 2.25x over-fetch) for tiles whose taps fit it: SLOWER
//     (64x180x240: 185.5 against 179.3 us; 8x480x640: 156.6 against 152.9) -- L2 -> SM traffic is not the bound; the
//     ncu source view puts the largest stall (17 %) on the wait for the stage to land;
//   * 64x16 tiles + a 3x3 menu of tensor-map boxes (over-fetch 1.3x instead of 2.25x), 2 CTAs per SM:
//     28 % / 42 % against 44 % / 57 % for this kernel at 180x240 / 480x640 -- occupancy (3 small CTAs,
//     2 pixels per thread) beats bytes saved.
// The optional image part (1 channel, full resolution) of the per-frame step rides in the same
// launch through the direct path.  Row pitches that are not a multiple of 16 bytes (W % 4 != 0) go
// through the quad-row tensor map (QUAD, below) when C % 32 == 0 and the grid is at least one full
// wave; everything else uses warp.cu.
#include <string.h>

#include "tma.cuh"
#include "warp_common.cuh"

namespace cf {

namespace wt {
constexpr int TW = 32, TH = 16;          // output tile
constexpr int BW = 48, BH = 24;          // source box (pixels)
constexpr int CC = 8;                    // channels per stage
constexpr int STAGES = 2;
constexpr int STAGE_FLOATS = CC * BH * BW;
constexpr int STAGE_BYTES = STAGE_FLOATS * 4;              // 36 864
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 64;  // + barriers
constexpr int CH_PER_CTA = 64;           // channel group of one CTA (8 stages of work).  32: the per-tile prologue (taps, bounding box: a third of
                                         // all instructions) is paid twice as often -- 64x180x240 180 us against 161 us (71 % of HBM), 8x480x640 154
                                         // against 141 us (73 %); 128: 162 / 145 us
}  // namespace wt

// One 32x32 block of the (few-channel, full-resolution) image: thread <-> 4 rows; all flow loads, then all 16 tap
// loads, then the stores (one pixel per thread cost two dependent DRAM round trips per 256 pixels).
__device__ __forceinline__ void warp_image_block(const WarpJob &ji, const float *__restrict__ flow, int fH, int fW, float sign,
                                                 bool identity, int b, int bx0, int by0, const FlowLR &lr) {
    const int Hi = ji.H, Wi = ji.W;
    const size_t hw = (size_t)Hi * Wi;
    const float *fbi = flow + (size_t)b * 2 * fH * fW;  // image resolution == flow resolution
    const int x = bx0 + (int)(threadIdx.x & 31);
    const int xc = min(x, Wi - 1);
    int yy[4];
    float fu[4], fv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        yy[k] = by0 + (int)(threadIdx.x >> 5) + 8 * k;
        const int yc = min(yy[k], Hi - 1);
        if (lr.lr) {   // fused upflow8 + unpad; the image part also writes the up-sampled flow out (once per pixel)
            const float2 uv = upflow8_at(lr, b, xc, yc);
            fu[k] = uv.x;
            fv[k] = uv.y;
            if (lr.flow_out != nullptr && x < Wi && yy[k] < Hi) {
                float *fo = lr.flow_out + (size_t)b * 2 * hw;
                st_cs(fo + (size_t)yc * Wi + xc, uv.x);
                st_cs(fo + hw + (size_t)yc * Wi + xc, uv.y);
            }
        } else {
            fu[k] = __ldg(fbi + (size_t)yc * Wi + xc);
            fv[k] = __ldg(fbi + hw + (size_t)yc * Wi + xc);
        }
    }
    Taps t[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        t[k] = identity ? identity_taps(xc, min(yy[k], Hi - 1), Wi) : make_taps(fu[k], fv[k], xc, min(yy[k], Hi - 1), Hi, Wi, sign);
    for (int c = 0; c < ji.C; ++c) {
        const float *src = ji.img + ((size_t)b * ji.C + c) * hw;
        float *dst = ji.out + ((size_t)b * ji.C + c) * hw;
        float v[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k][0] = __ldg(src + t[k].o00); v[k][1] = __ldg(src + t[k].o01);
            v[k][2] = __ldg(src + t[k].o10); v[k][3] = __ldg(src + t[k].o11);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float r = v[k][0] * t[k].w00;
            r += v[k][1] * t[k].w01;
            r += v[k][2] * t[k].w10;
            r += v[k][3] * t[k].w11;
            if (x < Wi && yy[k] < Hi) st_cs(dst + (size_t)yy[k] * Wi + x, r);
        }
    }
}

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *m, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(ptx::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ bool elect_lane() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}

// QUAD: tensors without a 16-byte row pitch (W % 4 != 0, e.g. the 130x173 codes of a 260x346 sensor).  No tensor map
// can describe their rows, but the whole buffer seen as a matrix of 4*W floats per row -- four source rows side by side,
// pitch 16*W bytes -- can: a [48 x 6] box at column (g % 4)*W + x0, map row g/4 is source rows g, g+4, .., g+20 (g = the
// global source-row index (b*C + c)*H + y).  Four boxes (g = g0 .. g0+3, g0 the row of the tile's first source row)
// bring a channel's 24 source rows, which land residue-major: row g0 + k is row k/4 of box k % 4.  But
//   * the box column must be a multiple of 4 floats (a start coordinate that is not 16-byte aligned raises "illegal
//     instruction" on the B200 -- scripts/experiments/quad_probe.cu), so a box starts at ((g % 4)*W + x0) & ~3 and its
//     data sits ((g % 4)*W + x0) % 4 floats further right: the usable box is 45 columns;
//   * that shift depends on q = g0 % 4, hence on the channel through c*H % 4.  A stage therefore holds 8 channels of
//     one residue class -- channel stride P = 1, 2 or 4 for H % 4 == 0, H even, H odd -- so that q is uniform per stage
//     and the tap offsets are recomputed per stage (a few integer operations), not per channel.
// Channels P apart are P*H/4 map rows apart -- a whole number -- so the map gets a third dimension (groups of P
// channels, stride 4*W*P*H bytes) and ONE [48 x 6 x 8] box per residue brings the stage's 8 channels: 4 instructions per
// stage (a 2-D map with one box per channel and residue, 32 instructions per stage spread over the 8 warps, measured
// the same: 64x260x346 399 against 405 us -- the instruction count is not the bound).
template <bool QUAD>
__global__ void __launch_bounds__(256, 3)
warp_tma_kernel(const __grid_constant__ CUtensorMap tmap, WarpJob ji, int image_in_tiles, WarpJob jz, int tiles_x, int tiles,
                int groups, const float *__restrict__ flow, int fH, int fW, float sign, const int *__restrict__ gate,
                const FlowLR lr) {
    using namespace wt;
    const bool identity = gate_closed(gate);  // device-side `not flow_final.any()` (e2v_model.py:184): copy
    const int b = blockIdx.y;
    int blk = blockIdx.x;
    const int n_codes_blocks = tiles * groups;
    // Image part.  Either CTAs of their own after the codes blocks in launch order (one 32x32 pixel block each), or --
    // image_in_tiles -- the first channel group's CTA of a codes tile (32x16 at half resolution) also warps the 64x32
    // full-resolution image pixels over that tile, before its own work: as CTAs of their own the image blocks are held
    // to this kernel's 3 CTAs per SM by its shared-memory footprint and run as a latency-bound tail (8x480x640: fused
    // 147 us against 130 us for the codes alone).  ONE copy of the block code serves both (a second and third inlined
    // copy cost 7 us per step at configs[1] through the instruction cache).
    const bool image_cta = blk >= n_codes_blocks;
    const int cblk = image_cta ? 0 : blk;
    const int group = cblk / tiles, tile = cblk - group * tiles;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    {
        int n_blocks = 0, bx0 = 0, by0 = 0;
        if (image_cta) {
            const int iblk = blk - n_codes_blocks, itx = (ji.W + 31) / 32;
            n_blocks = 1;
            by0 = (iblk / itx) * 32;
            bx0 = (iblk - (iblk / itx) * itx) * 32;
        } else if (image_in_tiles && group == 0) {
            n_blocks = 2;
            bx0 = 64 * tx;
            by0 = 32 * ty;
        }
#pragma unroll 1
        for (int k = 0; k < n_blocks; ++k) warp_image_block(ji, flow, fH, fW, sign, identity, b, bx0 + 32 * k, by0, lr);
        if (image_cta) return;
    }

    // declared aligned, no run-time rounding: keeps the pointers in the shared address space (LDS, not generic LD.E)
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float *stage0 = reinterpret_cast<float *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(stage0 + STAGES * STAGE_FLOATS);
    __shared__ int s_box[4];   // min x0, min y0, max x1, max y1
    __shared__ int red[4][8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = jz.H, W = jz.W;
    const float *fb = flow + (size_t)b * 2 * fH * fW;

    // ---- sample positions of this thread's two pixels.  A warp covers two tile rows as two passes of 16 px x 2 rows
    //      (lanes 0-15: row 2*warp, lanes 16-31: row 2*warp + 1; pass k: columns 16k .. 16k+15): consecutive box rows
    //      are 48 floats = 16 banks apart, so the 16 + 16 lanes of a pass start on 32 distinct banks (-4 % at
    //      64x180x240 against 32 lanes on one row; ncu still shows 1.67 wavefronts per LDS on a sheared flow)
    Taps taps[2];
    bool live[2];
    int x0a[2], y0a[2];
    int mnx = INT_MAX, mny = INT_MAX, mxx = -1, mxy = -1;
    // QUAD: a warp pass is 32 px of ONE row, pixel k on row 2*warp + k.  Output rows are only 4-byte aligned there: a
    // 64-byte half-warp store touches 3 sectors, a 128-byte warp store 5 (ncu at 32x260x346: 17.1 M -> 14.7 M store
    // sectors, 402 -> 384 us at 64 streams, although LDS wavefronts rose 25.4 M -> 28.5 M).  Source rows one apart sit
    // in different boxes a multiple of 32 banks apart, so the 16 x 2 layout has no conflict-free row pairing anyway
    const int xl = tx * TW + (QUAD ? lane : (lane & 15));
    const int yl = ty * TH + 2 * warp + (QUAD ? 0 : (lane >> 4));
    constexpr int KX = QUAD ? 0 : 16, KY = QUAD ? 1 : 0;   // pixel k of a thread: (xl + KX*k, yl + KY*k)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int x = xl + KX * k, y = yl + KY * k;
        live[k] = x < W && y < H;
        if (live[k]) {
            if (identity) {
                taps[k] = identity_taps(x, y, W);
            } else {
                const float2 uv = flow_at(fb, x, y, W, fH, fW, jz.half != 0, jz.sy, jz.sx, lr, b);
                taps[k] = make_taps(uv.x, uv.y, x, y, H, W, sign);
            }
            y0a[k] = taps[k].o00 / W;
            x0a[k] = taps[k].o00 - y0a[k] * W;
            mnx = min(mnx, x0a[k]); mny = min(mny, y0a[k]);
            mxx = max(mxx, taps[k].o01 - y0a[k] * W);          // x1 (clamped to W-1)
            mxy = max(mxy, taps[k].o10 / W);                   // y1 (clamped to H-1)
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if (lane == 0) { red[0][warp] = mnx; red[1][warp] = mny; red[2][warp] = mxx; red[3][warp] = mxy; }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) ptx::mbar_init(&full[s], 1);
        ptx::fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        int a = INT_MAX, c = INT_MAX, d = -1, e = -1;
        for (int k = 0; k < 8; ++k) { a = min(a, red[0][k]); c = min(c, red[1][k]); d = max(d, red[2][k]); e = max(e, red[3][k]); }
        s_box[0] = QUAD ? a : a & ~3;  // 16-byte aligned box origin: keeps the TMA requests sector-aligned
        s_box[1] = c; s_box[2] = d; s_box[3] = e;
    }
    __syncthreads();
    const int bx = s_box[0], by = s_box[1];
    const bool fits = s_box[2] >= 0 && (s_box[2] - bx) < (QUAD ? BW - 3 : BW) && (s_box[3] - by) < BH;

    const int c_begin = group * CH_PER_CTA, c_end = min(jz.C, c_begin + CH_PER_CTA);
    const size_t plane = (size_t)H * W;
    const float *img_b = jz.img + (size_t)b * jz.C * plane;
    float *out_b = jz.out + (size_t)b * jz.C * plane;

    if (!fits) {  // CTA-uniform: flow discontinuity inside the tile -> direct gather for this tile
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (live[k]) {
                const int p = (yl + KY * k) * W + xl + KX * k;
                for (int c0 = c_begin; c0 < c_end; c0 += 8) warp_pixel<8>(img_b, out_b, taps[k], p, c0, c_end, plane);
            }
        return;
    }

    // tap offsets inside one channel of a stage
    int s00[2], s01[2], s10[2], s11[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (live[k]) {
            const int x1 = taps[k].o01 - y0a[k] * W, y1 = taps[k].o10 / W;
            const int r0 = y0a[k] - by, r1 = y1 - by;
            if (QUAD) {  // row and column kept apart: the shared-memory row depends on the stage's q
                s00[k] = r0; s01[k] = x0a[k] - bx; s10[k] = r1; s11[k] = x1 - bx;
            } else {
                s00[k] = r0 * BW + (x0a[k] - bx);
                s01[k] = r0 * BW + (x1 - bx);
                s10[k] = r1 * BW + (x0a[k] - bx);
                s11[k] = r1 * BW + (x1 - bx);
            }
        }
    }

    constexpr int CSTR = QUAD ? (BH / 4) * BW : BH * BW;   // channel stride inside a stage
    const int nchunks = (c_end - c_begin + CC - 1) / CC;
    // QUAD: stage `chunk` holds channels first_channel(chunk) + P*i, i = 0..7 (the group's channel count is a multiple of
    // 8*P)
    const int P = QUAD ? ((H & 3) == 0 ? 1 : ((H & 1) ? 4 : 2)) : 1;
    auto first_channel = [&](int chunk) { return c_begin + (chunk % P) + P * CC * (chunk / P); };
    auto issue = [&](int chunk) {
        float *dst = stage0 + (chunk % STAGES) * STAGE_FLOATS;
        uint64_t *bar = &full[chunk % STAGES];
        if (QUAD) {
            if (warp == 0) {
                __syncwarp();
                if (elect_lane()) {
                    ptx::mbar_expect_tx(bar, STAGE_BYTES);
                    const int group = (b * jz.C + c_begin) / P + CC * (chunk / P);
                    const int r0 = (chunk % P) * H + by;   // row inside the group of P channels
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        tma_load_3d(dst + j * (CC * (BH / 4) * BW), &tmap, (((r0 + j) & 3) * W + bx) & ~3, (r0 + j) >> 2, group, bar);
                }
                __syncwarp();
            }
        } else if (tid == 0) {
            ptx::mbar_expect_tx(bar, STAGE_BYTES);
            ptx::tma_load_4d(dst, &tmap, bx, by, c_begin + chunk * CC, b, bar);
        }
    };
    for (int k = 0; k < STAGES - 1 && k < nchunks; ++k) issue(k);
    for (int k = 0; k < nchunks; ++k) {
        const int nxt = k + STAGES - 1;
        if (nxt < nchunks) issue(nxt);  // slot (nxt % STAGES) was drained at the end of iteration k-1
        ptx::mbar_wait(&full[k % STAGES], (uint32_t)((k / STAGES) & 1));
        const float *st = stage0 + (k % STAGES) * STAGE_FLOATS;
        const int c0 = QUAD ? first_channel(k) : c_begin + k * CC;
        const bool full_chunk = QUAD || c0 + CC <= c_end;
        const int q = QUAD ? (((b * jz.C + c0) * H + by) & 3) : 0;
        const size_t ostep = QUAD ? (size_t)P * plane : plane;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (!live[j]) continue;
            float *o = out_b + (size_t)c0 * plane + (size_t)(yl + KY * j) * W + xl + KX * j;
            int a0 = s00[j], a1 = s01[j], a2 = s10[j], a3 = s11[j];
            if (QUAD) {  // row g0 + k -> row k/4 of box k%4 ([8 channels][6 rows][48]), shifted right by (((g0 + k) % 4)*W + x0) % 4
                const int k0 = s00[j], k1 = s10[j];
                const int row0 = ((k0 & 3) * (CC * (BH / 4)) + (k0 >> 2)) * BW + ((((k0 + q) & 3) * W + bx) & 3);
                const int row1 = ((k1 & 3) * (CC * (BH / 4)) + (k1 >> 2)) * BW + ((((k1 + q) & 3) * W + bx) & 3);
                a0 = row0 + s01[j]; a1 = row0 + s11[j]; a2 = row1 + s01[j]; a3 = row1 + s11[j];
            }
            const float *s0 = st + a0, *s1 = st + a1, *s2 = st + a2, *s3 = st + a3;
            const float w0 = taps[j].w00, w1 = taps[j].w01, w2 = taps[j].w10, w3 = taps[j].w11;
            float v[CC][4];  // all 32 tap loads before the first store (LDS/STG interleaving serialises on aliasing)
#pragma unroll
            for (int c = 0; c < CC; ++c) {
                v[c][0] = s0[c * CSTR];
                v[c][1] = s1[c * CSTR];
                v[c][2] = s2[c * CSTR];
                v[c][3] = s3[c * CSTR];
            }
#pragma unroll
            for (int c = 0; c < CC; ++c) {
                float r = v[c][0] * w0;
                r += v[c][1] * w1;
                r += v[c][2] * w2;
                r += v[c][3] * w3;
                if (full_chunk || c0 + c < c_end) st_cs(o, r);
                o += ostep;
            }
        }
        __syncthreads();  // every thread is done with this stage before it is refilled
    }
}

int launch_warp_tma(const WarpJob &ji, bool with_image, const WarpJob &jz, const float *flow, int fH, int fW, float sign,
                    int B, const int *gate, const FlowLR &lr, cudaStream_t stream) {
    using namespace wt;
    static const char *env = getenv("CF_WARP_PATH");  // experiments: "direct" forces the plain gather
    const bool disabled = env && !strcmp(env, "direct");
    if (disabled || jz.C < CC || !aligned16(jz.img)) return 1;
    const bool quad = jz.W % 4 != 0;
    const int64_t src_rows = (int64_t)B * jz.C * jz.H;
    const char *env_quad = quad ? getenv("CF_WARP_QUAD") : nullptr;  // "0" keeps odd pitches on the direct path, "1" stages them at any size
    // one wave of CTAs or less is latency-bound and the four-box stages land later than the direct gather's taps
    // (1x260x346: 21.8 against 14.2 us)
    const int64_t quad_ctas = (int64_t)ceil_div(jz.W, TW) * ceil_div(jz.H, TH) * ceil_div(jz.C, CH_PER_CTA) * B;
    // (planes smaller than a box stay on the direct path too)
    if (quad && (jz.C % (4 * CC) != 0 || jz.W < 16 || jz.H < BH || (quad_ctas < 3 * (int64_t)sm_count() && !(env_quad && !strcmp(env_quad, "1"))) || src_rows > INT_MAX || (env_quad && !strcmp(env_quad, "0")))) return 1;
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return 1;
    CUtensorMap tmap;
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r;
    if (quad) {
        const int P = (jz.H & 3) == 0 ? 1 : ((jz.H & 1) ? 4 : 2);   // as in the kernel
        cuuint64_t dims[3] = {(cuuint64_t)jz.W * 4, (cuuint64_t)P * jz.H / 4, (cuuint64_t)(src_rows / ((int64_t)P * jz.H))};
        cuuint64_t strides[2] = {(cuuint64_t)jz.W * 16, (cuuint64_t)jz.W * 4 * P * jz.H};
        cuuint32_t box[3] = {BW, BH / 4, CC};
        r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(jz.img), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[4] = {(cuuint64_t)jz.W, (cuuint64_t)jz.H, (cuuint64_t)jz.C, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)jz.W * 4, (cuuint64_t)jz.W * jz.H * 4, (cuuint64_t)jz.W * jz.H * jz.C * 4};
        cuuint32_t box[4] = {BW, BH, CC, 1};
        r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(jz.img), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    CF_REQUIRE(r == CUDA_SUCCESS, CF_ERR_CUDA, "cuTensorMapEncodeTiled (warp) failed with CUresult %d", (int)r);
    int dev = 0;
    CF_CUDA(cudaGetDevice(&dev));
    static bool opt_in[64] = {};
    if (!opt_in[dev & 63]) {
        CF_CUDA(cudaFuncSetAttribute(warp_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CF_CUDA(cudaFuncSetAttribute(warp_tma_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CF_CUDA(cudaFuncSetAttribute(warp_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CF_CUDA(cudaFuncSetAttribute(warp_tma_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        opt_in[dev & 63] = true;
    }
    const int tiles_x = (int)ceil_div(jz.W, TW), tiles = tiles_x * (int)ceil_div(jz.H, TH);
    const int groups = (int)ceil_div(jz.C, CH_PER_CTA);
    // even image sizes (the reference requires them, data_readers/video_readers.py:403-404): the 64x32 image pixels over
    // a codes tile are exactly that tile's share; odd sizes keep separate image CTAs
    static const char *env_img = getenv("CF_WARP_IMAGE_CTAS");   // experiments: "1" forces separate image CTAs
    // ... and only when the codes tiles run in several waves (64x180x240: 160 -> 156 us, 8x480x640: 141 -> 137 us): in a
    // single wave the longer first-group CTAs are the critical path (8x180x240: 27.7 against 26.5 us)
    const int64_t code_ctas = (int64_t)ceil_div(jz.W, TW) * ceil_div(jz.H, TH) * ceil_div(jz.C, CH_PER_CTA) * B;
    const bool in_tiles = with_image && ji.H == 2 * jz.H && ji.W == 2 * jz.W && code_ctas > 6 * (int64_t)sm_count() &&
                          !(env_img && !strcmp(env_img, "1"));
    const int n_img = (with_image && !in_tiles) ? (int)(ceil_div(ji.W, 32) * ceil_div(ji.H, 32)) : 0;
    dim3 grid((unsigned)(n_img + tiles * groups), (unsigned)B);
    if (quad)
        warp_tma_kernel<true><<<grid, 256, SMEM_BYTES, stream>>>(tmap, ji, in_tiles ? 1 : 0, jz, tiles_x, tiles, groups, flow, fH, fW, sign, gate, lr);
    else
        warp_tma_kernel<false><<<grid, 256, SMEM_BYTES, stream>>>(tmap, ji, in_tiles ? 1 : 0, jz, tiles_x, tiles, groups, flow, fH, fW, sign, gate, lr);
    CF_LAUNCH_CHECK("warp_tma_kernel");
    return CF_OK;
}

}  // namespace cf
