// Flow-guided bilinear warp (part 2 of the hot path).
//
// Replaces forwardWarp / backWarp / FrameWarp of the reference
// (utils/flow_utils.py:83-120, 153-190, 212-221).  Both reference "modes" are a
// bilinear GATHER through grid_sample(align_corners=True, padding_mode=
// 'reflection') at (x -/+ u, y -/+ v) normalised as 2*(x/W - 0.5); they differ
// only in the sign of the flow.
//
// Data layout in HBM: img/out [B,C,H,W], flow [B,2,fH,fW], all fp32 row-major.
// Roofline: HBM.  Algorithmic bytes per call = B*H*W*(8*C + 8): every input
// and output channel once plus the two flow channels.  The gather has high
// spatial locality (network flow is smooth: it is predicted at 1/8 resolution
// and up-sampled), so the 4 taps of neighbouring threads fall into the same
// 128-byte lines.  One CTA owns a 32x8 pixel tile (the two tap rows of a pixel
// row are re-used by the row below out of L1), one thread owns one output
// pixel, computes the sample position once and streams its channel group in
// batches of CPT channels with all 4*CPT loads in flight before the first FMA.
// No fast-math in this file: the position arithmetic mirrors ATen's operation
// order (true fp32 division) so that floor() picks the same taps.
#include "common.cuh"

namespace cf {

struct Taps {
    int o00, o01, o10, o11;  // offsets inside one channel plane
    float w00, w01, w10, w11;
};

// ATen reflect_coordinates(v, 0, 2*(size-1)) followed by clip_coordinates.
__device__ __forceinline__ float reflect_clip(float v, int size) {
    if (size == 1) return 0.f;
    const float span = (float)(size - 1);
    const float a = fabsf(v);
    const float extra = fmodf(a, span);
    const int flips = (int)floorf(a / span);
    const float r = (flips & 1) ? span - extra : extra;
    return fminf(span, fmaxf(r, 0.f));
}

// Flow at output pixel (x, y).  half == false: flow has the output's size.
// half == true: x0.5 bilinear, align_corners=True from the [fH, fW] field
// (ATen upsample_bilinear2d: src = scale*dst, lambda clamped to [0,1]).
__device__ __forceinline__ float2 flow_at(const float *__restrict__ fb, int x, int y, int W,
                                          int fH, int fW, bool half, float sy, float sx) {
    if (!half) {
        const int p = y * W + x;
        return make_float2(__ldg(fb + p), __ldg(fb + (size_t)fH * fW + p));
    }
    const float fy = sy * (float)y, fx = sx * (float)x;
    int y0 = min((int)fy, fH - 1), x0 = min((int)fx, fW - 1);
    const int y1 = y0 + (y0 < fH - 1 ? 1 : 0), x1 = x0 + (x0 < fW - 1 ? 1 : 0);
    const float ly1 = fminf(fmaxf(fy - (float)y0, 0.f), 1.f), lx1 = fminf(fmaxf(fx - (float)x0, 0.f), 1.f);
    const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    float2 r;
    const float *c = fb;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float p00 = __ldg(c + y0 * fW + x0), p01 = __ldg(c + y0 * fW + x1);
        const float p10 = __ldg(c + y1 * fW + x0), p11 = __ldg(c + y1 * fW + x1);
        const float v = ly0 * (lx0 * p00 + lx1 * p01) + ly1 * (lx0 * p10 + lx1 * p11);
        if (k == 0) r.x = v; else r.y = v;
        c += (size_t)fH * fW;
    }
    return r;
}

__device__ __forceinline__ Taps make_taps(float u, float v, int x, int y, int H, int W, float sign) {
    // utils/flow_utils.py:110-116 (sign=+1) / :180-186 (sign=-1), then ATen
    // grid_sampler_unnormalize(align_corners=True): ((g + 1) / 2) * (size - 1).
    const float gx = 2.f * (((float)x + sign * u) / (float)W - 0.5f);
    const float gy = 2.f * (((float)y + sign * v) / (float)H - 0.5f);
    const float ix = reflect_clip(((gx + 1.f) / 2.f) * (float)(W - 1), W);
    const float iy = reflect_clip(((gy + 1.f) / 2.f) * (float)(H - 1), H);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    // after the clip x0+1 == W only when ix == W-1 exactly, where its weight is 0
    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    // ATen weights: nw = (x_se - ix)(y_se - iy), ne = (ix - x_sw)(y_sw - iy), ...
    const float ax1 = (fx0 + 1.f) - ix, ax0 = ix - fx0;
    const float ay1 = (fy0 + 1.f) - iy, ay0 = iy - fy0;
    Taps t;
    t.o00 = y0 * W + x0; t.o01 = y0 * W + x1; t.o10 = y1 * W + x0; t.o11 = y1 * W + x1;
    t.w00 = ax1 * ay1;
    t.w01 = ax0 * ay1;
    t.w10 = ax1 * ay0;
    t.w11 = ax0 * ay0;
    return t;
}

// One thread = one output pixel x CPT channels.
template <int CPT>
__device__ __forceinline__ void warp_pixel(const float *__restrict__ img_b, float *__restrict__ out_b,
                                           const Taps &t, int p, int c0, int C, size_t plane) {
    float v[CPT][4];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int c = c0 + k;
        if (c < C) {
            const float *s = img_b + (size_t)c * plane;
            v[k][0] = __ldg(s + t.o00); v[k][1] = __ldg(s + t.o01);
            v[k][2] = __ldg(s + t.o10); v[k][3] = __ldg(s + t.o11);
        }
    }
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int c = c0 + k;
        if (c < C) {
            // ATen order: nw, ne, sw, se accumulated left to right
            float r = v[k][0] * t.w00;
            r += v[k][1] * t.w01;
            r += v[k][2] * t.w10;
            r += v[k][3] * t.w11;
            st_cs(out_b + (size_t)c * plane + p, r);
        }
    }
}

struct WarpJob {
    const float *img; float *out;
    int C, H, W;          // geometry of img/out
    int half;             // 1: flow is [2, fH, fW] at twice the resolution
    float sy, sx;         // align_corners scales for the fused down-sampling
    int tiles_x;          // 32-pixel-wide tiles per row
    int blocks_x;         // pixel tiles (32x8) per (batch, channel group)
    int cpg;              // channels per group (one thread loops over them, CPT at a time)
    int groups;           // channel groups
};
constexpr int kTileW = 32, kTileH = 8;

template <int CPT>
__device__ __forceinline__ void run_job(const WarpJob &j, const float *__restrict__ flow,
                                        int fH, int fW, float sign, int tile, int group, int b) {
    const int ty = tile / j.tiles_x, tx = tile - ty * j.tiles_x;
    const int x = tx * kTileW + (threadIdx.x & 31), y = ty * kTileH + (threadIdx.x >> 5);
    if (x >= j.W || y >= j.H) return;
    const int p = y * j.W + x;
    const float *fb = flow + (size_t)b * 2 * fH * fW;
    const float2 uv = flow_at(fb, x, y, j.W, fH, fW, j.half != 0, j.sy, j.sx);
    const Taps t = make_taps(uv.x, uv.y, x, y, j.H, j.W, sign);
    const size_t plane = (size_t)j.H * j.W;
    const float *img_b = j.img + (size_t)b * j.C * plane;
    float *out_b = j.out + (size_t)b * j.C * plane;
    const int c_end = min(j.C, (group + 1) * j.cpg);
    for (int c0 = group * j.cpg; c0 < c_end; c0 += CPT) warp_pixel<CPT>(img_b, out_b, t, p, c0, c_end, plane);
}

template <int CPT>
__global__ void __launch_bounds__(256, 4) warp_gather_kernel(WarpJob j, const float *__restrict__ flow,
                                                          int fH, int fW, float sign) {
    run_job<CPT>(j, flow, fH, fW, sign, blockIdx.x, blockIdx.y, blockIdx.z);
}

// image (CPT=1 per thread, few channels) + codes (CPT=8, 32 channels per thread) in one launch:
// blockIdx.x < img.blocks_x*img.groups -> image part, the rest -> codes part.
__global__ void __launch_bounds__(256, 4) warp_frame_and_codes_kernel(WarpJob ji, WarpJob jz,
                                                                   const float *__restrict__ flow,
                                                                   int fH, int fW, float sign) {
    const int b = blockIdx.y;
    int blk = blockIdx.x;
    const int n_img = ji.blocks_x * ji.groups;
    if (blk < n_img) {
        run_job<1>(ji, flow, fH, fW, sign, blk % ji.blocks_x, blk / ji.blocks_x, b);
    } else {
        blk -= n_img;
        run_job<8>(jz, flow, fH, fW, sign, blk % jz.blocks_x, blk / jz.blocks_x, b);
    }
}

static int make_job(WarpJob &j, const float *img, float *out, int C, int H, int W, int fH, int fW, int cpt) {
    j.img = img; j.out = out; j.C = C; j.H = H; j.W = W;
    if (fH == H && fW == W) {
        j.half = 0; j.sy = j.sx = 0.f;
    } else if (H == fH / 2 && W == fW / 2) {
        j.half = 1;
        j.sy = H > 1 ? (float)(fH - 1) / (float)(H - 1) : 0.f;
        j.sx = W > 1 ? (float)(fW - 1) / (float)(W - 1) : 0.f;
    } else {
        set_error("cf_warp: flow is %dx%d but the image is %dx%d (must be equal, or image == flow/2)", fH, fW, H, W);
        return CF_ERR_INVALID_ARG;
    }
    j.tiles_x = (int)ceil_div(W, kTileW);
    j.blocks_x = j.tiles_x * (int)ceil_div(H, kTileH);
    j.cpg = C >= 4 * cpt ? 4 * cpt : (int)ceil_div(C, cpt) * cpt;  // up to 4 batches of CPT channels per thread
    j.groups = (int)ceil_div(C, j.cpg);
    return CF_OK;
}

}  // namespace cf

extern "C" int cf_warp(const float *img, const float *flow, float *out, int B, int C, int H, int W,
                       int flowH, int flowW, float sign, cf_stream_t stream_) {
    using namespace cf;
    if (int rc = check_device()) return rc;
    CF_REQUIRE(img && flow && out, CF_ERR_NULL, "cf_warp: null pointer");
    CF_REQUIRE(B >= 0 && C >= 0 && H > 0 && W > 0 && flowH > 0 && flowW > 0, CF_ERR_INVALID_ARG,
               "cf_warp: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
    CF_REQUIRE(sign == 1.f || sign == -1.f, CF_ERR_INVALID_ARG, "cf_warp: sign must be +1 (backward) or -1 (forward)");
    CF_REQUIRE((int64_t)H * W < (1ll << 30), CF_ERR_INVALID_ARG, "cf_warp: plane too large");
    CF_REQUIRE(B <= 65535, CF_ERR_INVALID_ARG, "cf_warp: B > 65535");
    if (B == 0 || C == 0) return CF_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int cpt = C >= 8 ? 8 : (C >= 4 ? 4 : 1);
    WarpJob j;
    if (int rc = make_job(j, img, out, C, H, W, flowH, flowW, cpt)) return rc;
    CF_REQUIRE(j.groups <= 65535, CF_ERR_INVALID_ARG, "cf_warp: too many channels");
    dim3 grid(j.blocks_x, j.groups, B);
    if (cpt == 8) warp_gather_kernel<8><<<grid, 256, 0, stream>>>(j, flow, flowH, flowW, sign);
    else if (cpt == 4) warp_gather_kernel<4><<<grid, 256, 0, stream>>>(j, flow, flowH, flowW, sign);
    else warp_gather_kernel<1><<<grid, 256, 0, stream>>>(j, flow, flowH, flowW, sign);
    CF_LAUNCH_CHECK("warp_gather_kernel");
    return CF_OK;
}

extern "C" int cf_warp_frame_and_codes(const float *img, const float *codes, const float *flow,
                                       float *img_out, float *codes_out, int B, int Ci, int Cz,
                                       int H, int W, float sign, cf_stream_t stream_) {
    using namespace cf;
    if (int rc = check_device()) return rc;
    CF_REQUIRE(img && codes && flow && img_out && codes_out, CF_ERR_NULL, "cf_warp_frame_and_codes: null pointer");
    CF_REQUIRE(B >= 0 && Ci > 0 && Cz > 0 && H > 1 && W > 1, CF_ERR_INVALID_ARG,
               "cf_warp_frame_and_codes: bad shape B=%d Ci=%d Cz=%d H=%d W=%d", B, Ci, Cz, H, W);
    CF_REQUIRE(sign == 1.f || sign == -1.f, CF_ERR_INVALID_ARG, "cf_warp_frame_and_codes: sign must be +-1");
    CF_REQUIRE((int64_t)H * W < (1ll << 30) && B <= 65535, CF_ERR_INVALID_ARG, "cf_warp_frame_and_codes: too large");
    if (B == 0) return CF_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    WarpJob ji, jz;
    if (int rc = make_job(ji, img, img_out, Ci, H, W, H, W, 1)) return rc;
    if (int rc = make_job(jz, codes, codes_out, Cz, H / 2, W / 2, H, W, 8)) return rc;
    dim3 grid(ji.blocks_x * ji.groups + jz.blocks_x * jz.groups, B);
    warp_frame_and_codes_kernel<<<grid, 256, 0, stream>>>(ji, jz, flow, H, W, sign);
    CF_LAUNCH_CHECK("warp_frame_and_codes_kernel");
    return CF_OK;
}
