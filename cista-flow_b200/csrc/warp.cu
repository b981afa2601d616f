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
#include "warp_common.cuh"

namespace cf {

// implemented in warp_tma.cu: CF_OK / error, or 1 when the TMA-staged path does not apply
int launch_warp_tma(const WarpJob &ji, bool with_image, const WarpJob &jz, const float *flow, int fH, int fW, float sign,
                    int B, const int *gate, const FlowLR &lr, cudaStream_t stream);

static int launch_staged(const WarpJob &ji, bool with_image, const WarpJob &jz, const float *flow, int fH, int fW,
                         float sign, int B, const int *gate, const FlowLR &lr, cudaStream_t stream) {
    return launch_warp_tma(ji, with_image, jz, flow, fH, fW, sign, B, gate, lr, stream);
}

// flag[0] = 1 iff any element of flow is non-zero (NaN counts, -0.0 does not: torch.Tensor.any()).  flag is zeroed
// by the host side before the launch; CTAs that see a non-zero element store 1 (same value from everyone: no atomic).
__global__ void __launch_bounds__(256) flow_any_kernel(const float *__restrict__ flow, int64_t n, int *__restrict__ flag) {
    bool any = false;
    const int64_t n4 = n >> 2;
    const float4 *f4 = reinterpret_cast<const float4 *>(flow);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(f4 + i);
        any = any || !(v.x == 0.f) || !(v.y == 0.f) || !(v.z == 0.f) || !(v.w == 0.f);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) any = any || !(flow[(n4 << 2) + threadIdx.x] == 0.f);
    if (__syncthreads_or(any) && threadIdx.x == 0) *flag = 1;
}

template <int CPT>
__global__ void __launch_bounds__(256, 4) warp_gather_kernel(WarpJob j, const float *__restrict__ flow,
                                                          int fH, int fW, float sign) {
    run_job<CPT>(j, flow, fH, fW, sign, blockIdx.x, blockIdx.y, blockIdx.z, nullptr, FlowLR{});
}

// image (CPT=1 per thread, few channels) + codes (CPT=8, 32 channels per thread) in one launch:
// blockIdx.x < img.blocks_x*img.groups -> image part, the rest -> codes part.
__global__ void __launch_bounds__(256, 4) warp_frame_and_codes_kernel(WarpJob ji, WarpJob jz,
                                                                   const float *__restrict__ flow,
                                                                   int fH, int fW, float sign, const int *__restrict__ gate,
                                                                   FlowLR lr) {
    const int b = blockIdx.y;
    int blk = blockIdx.x;
    const int n_img = ji.blocks_x * ji.groups;
    if (blk < n_img) {
        run_job<1>(ji, flow, fH, fW, sign, blk % ji.blocks_x, blk / ji.blocks_x, b, gate, lr, true);
    } else {
        blk -= n_img;
        run_job<8>(jz, flow, fH, fW, sign, blk % jz.blocks_x, blk / jz.blocks_x, b, gate, lr);
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
    if (C >= 8) {  // multi-channel tensors: TMA-staged kernel when the shape allows it
        const int rc = launch_staged(j, false, j, flow, flowH, flowW, sign, B, nullptr, FlowLR{}, stream);
        if (rc != 1) return rc;
    }
    dim3 grid(j.blocks_x, j.groups, B);
    if (cpt == 8) warp_gather_kernel<8><<<grid, 256, 0, stream>>>(j, flow, flowH, flowW, sign);
    else if (cpt == 4) warp_gather_kernel<4><<<grid, 256, 0, stream>>>(j, flow, flowH, flowW, sign);
    else warp_gather_kernel<1><<<grid, 256, 0, stream>>>(j, flow, flowH, flowW, sign);
    CF_LAUNCH_CHECK("warp_gather_kernel");
    return CF_OK;
}

extern "C" int cf_flow_any(const float *flow, int64_t n, int *flag, cf_stream_t stream_) {
    using namespace cf;
    if (int rc = check_device()) return rc;
    CF_REQUIRE(flag && (flow || n == 0), CF_ERR_NULL, "cf_flow_any: null pointer");
    CF_REQUIRE(n >= 0, CF_ERR_INVALID_ARG, "cf_flow_any: negative size");
    CF_REQUIRE(n == 0 || aligned16(flow), CF_ERR_ALIGN, "cf_flow_any: flow not 16-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    CF_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), stream));
    if (n == 0) return CF_OK;
    int64_t blocks = ceil_div(n, 256 * 4 * 4);  // ~4 float4 per thread
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    flow_any_kernel<<<(unsigned)blocks, 256, 0, stream>>>(flow, n, flag);
    CF_LAUNCH_CHECK("flow_any_kernel");
    return CF_OK;
}

extern "C" int cf_warp_frame_and_codes(const float *img, const float *codes, const float *flow,
                                       float *img_out, float *codes_out, int B, int Ci, int Cz,
                                       int H, int W, float sign, cf_stream_t stream_) {
    return cf_warp_frame_and_codes_gated(img, codes, flow, img_out, codes_out, B, Ci, Cz, H, W, sign, nullptr, stream_);
}

namespace cf {
static int warp_frame_and_codes_impl(const float *img, const float *codes, const float *flow, float *img_out, float *codes_out,
                                     int B, int Ci, int Cz, int H, int W, float sign, const int *gate, const FlowLR &lr,
                                     cudaStream_t stream) {
    WarpJob ji, jz;
    if (int rc = make_job(ji, img, img_out, Ci, H, W, H, W, 1)) return rc;
    if (int rc = make_job(jz, codes, codes_out, Cz, H / 2, W / 2, H, W, 8)) return rc;
    {
        const int rc = launch_staged(ji, true, jz, flow, H, W, sign, B, gate, lr, stream);
        if (rc != 1) return rc;
    }
    dim3 grid(ji.blocks_x * ji.groups + jz.blocks_x * jz.groups, B);
    warp_frame_and_codes_kernel<<<grid, 256, 0, stream>>>(ji, jz, flow, H, W, sign, gate, lr);
    CF_LAUNCH_CHECK("warp_frame_and_codes_kernel");
    return CF_OK;
}
}  // namespace cf

extern "C" int cf_warp_frame_and_codes_gated(const float *img, const float *codes, const float *flow,
                                             float *img_out, float *codes_out, int B, int Ci, int Cz,
                                             int H, int W, float sign, const int *gate, cf_stream_t stream_) {
    using namespace cf;
    if (int rc = check_device()) return rc;
    CF_REQUIRE(img && codes && flow && img_out && codes_out, CF_ERR_NULL, "cf_warp_frame_and_codes: null pointer");
    CF_REQUIRE(B >= 0 && Ci > 0 && Cz > 0 && H > 1 && W > 1, CF_ERR_INVALID_ARG,
               "cf_warp_frame_and_codes: bad shape B=%d Ci=%d Cz=%d H=%d W=%d", B, Ci, Cz, H, W);
    CF_REQUIRE(sign == 1.f || sign == -1.f, CF_ERR_INVALID_ARG, "cf_warp_frame_and_codes: sign must be +-1");
    CF_REQUIRE((int64_t)H * W < (1ll << 30) && B <= 65535, CF_ERR_INVALID_ARG, "cf_warp_frame_and_codes: too large");
    if (B == 0) return CF_OK;
    return warp_frame_and_codes_impl(img, codes, flow, img_out, codes_out, B, Ci, Cz, H, W, sign, gate, FlowLR{},
                                     (cudaStream_t)stream_);
}

extern "C" int cf_warp_frame_and_codes_upflow8(const float *img, const float *codes, const float *flow_lr, float *img_out,
                                               float *codes_out, float *flow_out, int B, int Ci, int Cz, int H, int W,
                                               int lh, int lw, int pad_h, int pad_w, float sign, cf_stream_t stream_) {
    using namespace cf;
    if (int rc = check_device()) return rc;
    CF_REQUIRE(img && codes && flow_lr && img_out && codes_out, CF_ERR_NULL, "cf_warp_frame_and_codes_upflow8: null pointer");
    CF_REQUIRE(B >= 0 && Ci > 0 && Cz > 0 && H > 1 && W > 1 && lh > 0 && lw > 0, CF_ERR_INVALID_ARG,
               "cf_warp_frame_and_codes_upflow8: bad shape B=%d Ci=%d Cz=%d H=%d W=%d lh=%d lw=%d", B, Ci, Cz, H, W, lh, lw);
    CF_REQUIRE(pad_h >= 0 && pad_w >= 0 && 8 * lh - pad_h == H && 8 * lw - pad_w == W, CF_ERR_INVALID_ARG,
               "cf_warp_frame_and_codes_upflow8: 8*[%d,%d] minus the top/left padding [%d,%d] is not the frame [%d,%d]", lh, lw,
               pad_h, pad_w, H, W);
    CF_REQUIRE(sign == 1.f || sign == -1.f, CF_ERR_INVALID_ARG, "cf_warp_frame_and_codes_upflow8: sign must be +-1");
    CF_REQUIRE((int64_t)H * W < (1ll << 30) && B <= 65535, CF_ERR_INVALID_ARG, "cf_warp_frame_and_codes_upflow8: too large");
    if (B == 0) return CF_OK;
    FlowLR lr;
    lr.lr = flow_lr; lr.lh = lh; lr.lw = lw; lr.pad_h = pad_h; lr.pad_w = pad_w;
    lr.sy = 8 * lh > 1 ? (float)(lh - 1) / (float)(8 * lh - 1) : 0.f;   // ATen area_pixel_compute_scale, align_corners
    lr.sx = 8 * lw > 1 ? (float)(lw - 1) / (float)(8 * lw - 1) : 0.f;
    lr.flow_out = flow_out;
    // (no zero-flow gate here: the reference's predicate is on the up-sampled flow, which this call never materialises
    //  before warping; callers that need the branch use cf_flow_any on flow_out of the previous step or the unfused calls)
    return warp_frame_and_codes_impl(img, codes, flow_lr /*unused*/, img_out, codes_out, B, Ci, Cz, H, W, sign, nullptr, lr,
                                     (cudaStream_t)stream_);
}
