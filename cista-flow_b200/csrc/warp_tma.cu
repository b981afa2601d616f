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
// The optional image part (1 channel, full resolution) of the per-frame step rides in the same
// launch through the direct path.  Needs a 16-byte aligned row pitch (W % 4 == 0) for the tensor
// map; other shapes use warp.cu.
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
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 128 + 64;  // + alignment slack + barriers
constexpr int CH_PER_CTA = 32;           // channel group of one CTA (4 stages of work)
}  // namespace wt

__global__ void __launch_bounds__(256, 3)
warp_tma_kernel(const __grid_constant__ CUtensorMap tmap, WarpJob ji, int n_img_blocks, WarpJob jz, int tiles_x, int tiles,
                int groups, const float *__restrict__ flow, int fH, int fW, float sign) {
    using namespace wt;
    const int b = blockIdx.y;
    int blk = blockIdx.x;
    if (blk < n_img_blocks) {  // image part: direct gather
        run_job<1>(ji, flow, fH, fW, sign, blk % ji.blocks_x, blk / ji.blocks_x, b);
        return;
    }
    blk -= n_img_blocks;
    const int group = blk / tiles, tile = blk - group * tiles;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;

    extern __shared__ uint8_t smem_raw[];
    float *stage0 = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t *full = reinterpret_cast<uint64_t *>(stage0 + STAGES * STAGE_FLOATS);
    __shared__ int s_box[4];   // min x0, min y0, max x1, max y1
    __shared__ int red[4][8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = jz.H, W = jz.W;
    const float *fb = flow + (size_t)b * 2 * fH * fW;

    // ---- sample positions of this thread's two pixels (rows warp and warp + 8 of the tile)
    Taps taps[2];
    bool live[2];
    int x0a[2], y0a[2];
    int mnx = INT_MAX, mny = INT_MAX, mxx = -1, mxy = -1;
    const int x = tx * TW + lane;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int y = ty * TH + warp + 8 * k;
        live[k] = x < W && y < H;
        if (live[k]) {
            const float2 uv = flow_at(fb, x, y, W, fH, fW, jz.half != 0, jz.sy, jz.sx);
            taps[k] = make_taps(uv.x, uv.y, x, y, H, W, sign);
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
        s_box[0] = a & ~3;  // 16-byte aligned box origin: keeps the TMA requests sector-aligned
        s_box[1] = c; s_box[2] = d; s_box[3] = e;
    }
    __syncthreads();
    const int bx = s_box[0], by = s_box[1];
    const bool fits = s_box[2] >= 0 && (s_box[2] - bx) < BW && (s_box[3] - by) < BH;

    const int c_begin = group * CH_PER_CTA, c_end = min(jz.C, c_begin + CH_PER_CTA);
    const size_t plane = (size_t)H * W;
    const float *img_b = jz.img + (size_t)b * jz.C * plane;
    float *out_b = jz.out + (size_t)b * jz.C * plane;

    if (!fits) {  // CTA-uniform: flow discontinuity inside the tile -> direct gather for this tile
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (live[k]) {
                const int p = (ty * TH + warp + 8 * k) * W + x;
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
            s00[k] = (y0a[k] - by) * BW + (x0a[k] - bx);
            s01[k] = (y0a[k] - by) * BW + (x1 - bx);
            s10[k] = (y1 - by) * BW + (x0a[k] - bx);
            s11[k] = (y1 - by) * BW + (x1 - bx);
        }
    }

    const int nchunks = (c_end - c_begin + CC - 1) / CC;
    if (tid == 0) {
        for (int k = 0; k < STAGES - 1 && k < nchunks; ++k) {
            ptx::mbar_expect_tx(&full[k], STAGE_BYTES);
            ptx::tma_load_4d(stage0 + k * STAGE_FLOATS, &tmap, bx, by, c_begin + k * CC, b, &full[k]);
        }
    }
    for (int k = 0; k < nchunks; ++k) {
        const int nxt = k + STAGES - 1;
        if (tid == 0 && nxt < nchunks) {  // slot (nxt % STAGES) was drained at the end of iteration k-1
            ptx::mbar_expect_tx(&full[nxt % STAGES], STAGE_BYTES);
            ptx::tma_load_4d(stage0 + (nxt % STAGES) * STAGE_FLOATS, &tmap, bx, by, c_begin + nxt * CC, b, &full[nxt % STAGES]);
        }
        ptx::mbar_wait(&full[k % STAGES], (uint32_t)((k / STAGES) & 1));
        const float *st = stage0 + (k % STAGES) * STAGE_FLOATS;
        const int c0 = c_begin + k * CC;
        const bool full_chunk = c0 + CC <= c_end;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (!live[j]) continue;
            float *o = out_b + (size_t)c0 * plane + (size_t)(ty * TH + warp + 8 * j) * W + x;
            const float *s0 = st + s00[j], *s1 = st + s01[j], *s2 = st + s10[j], *s3 = st + s11[j];
            const float w0 = taps[j].w00, w1 = taps[j].w01, w2 = taps[j].w10, w3 = taps[j].w11;
            if (full_chunk) {
#pragma unroll
                for (int c = 0; c < CC; ++c) {
                    float r = s0[c * (BH * BW)] * w0;
                    r += s1[c * (BH * BW)] * w1;
                    r += s2[c * (BH * BW)] * w2;
                    r += s3[c * (BH * BW)] * w3;
                    st_cs(o, r);
                    o += plane;
                }
            } else {
                for (int c = 0; c0 + c < c_end; ++c) {
                    float r = s0[c * (BH * BW)] * w0;
                    r += s1[c * (BH * BW)] * w1;
                    r += s2[c * (BH * BW)] * w2;
                    r += s3[c * (BH * BW)] * w3;
                    st_cs(o, r);
                    o += plane;
                }
            }
        }
        __syncthreads();  // every thread is done with this stage before it is refilled
    }
}

int launch_warp_tma(const WarpJob &ji, bool with_image, const WarpJob &jz, const float *flow, int fH, int fW, float sign,
                    int B, cudaStream_t stream) {
    using namespace wt;
    static const bool disabled = getenv("CF_WARP_NO_TMA") != nullptr;  // experiments: force the direct gather
    if (disabled || jz.W % 4 != 0 || jz.C < CC || !aligned16(jz.img)) return 1;
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return 1;
    CUtensorMap tmap;
    cuuint64_t dims[4] = {(cuuint64_t)jz.W, (cuuint64_t)jz.H, (cuuint64_t)jz.C, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)jz.W * 4, (cuuint64_t)jz.W * jz.H * 4, (cuuint64_t)jz.W * jz.H * jz.C * 4};
    cuuint32_t box[4] = {BW, BH, CC, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(jz.img), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CF_REQUIRE(r == CUDA_SUCCESS, CF_ERR_CUDA, "cuTensorMapEncodeTiled (warp) failed with CUresult %d", (int)r);
    int dev = 0;
    CF_CUDA(cudaGetDevice(&dev));
    static bool opt_in[64] = {};
    if (!opt_in[dev & 63]) {
        CF_CUDA(cudaFuncSetAttribute(warp_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        opt_in[dev & 63] = true;
    }
    const int tiles_x = (int)ceil_div(jz.W, TW), tiles = tiles_x * (int)ceil_div(jz.H, TH);
    const int groups = (int)ceil_div(jz.C, CH_PER_CTA);
    const int n_img = with_image ? ji.blocks_x * ji.groups : 0;
    dim3 grid((unsigned)(n_img + tiles * groups), (unsigned)B);
    warp_tma_kernel<<<grid, 256, SMEM_BYTES, stream>>>(tmap, ji, n_img, jz, tiles_x, tiles, groups, flow, fH, fW, sign);
    CF_LAUNCH_CHECK("warp_tma_kernel");
    return CF_OK;
}

}  // namespace cf
