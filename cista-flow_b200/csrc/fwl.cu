// Flow-warp of an event voxel grid for the FWL metric.
//
// Replaces voxel_warping_flow_loss (loss.py:27-83 of the reference; call sites test_wo_flow.py:161,
// test_mvsec.py:180).  Channel i of the voxel grid is sampled bilinearly (zeros padding,
// align_corners=True) at (x + dx * r_i, y + dy * r_i), r_i = i / (C - 1) -- or 1 - i / (C - 1) with the
// displacement negated for reverse_time -- through the reference's normalisation 2 * coord / size - 1,
// which under align_corners=True samples at coord * (size - 1) / size (the same quirk as
// utils/flow_utils.py).  The C warped channels are summed into one image and the metric is
// the unbiased variance of that image over the whole batch (tensor.var()).
//
// The reference runs C grid_sample calls over the WHOLE C-channel grid and keeps one channel of each
// (C^2 * H * W samples, C full passes); here one thread owns one pixel, reads the displacement once and
// visits every channel once: algorithmic bytes = 4 * B*H*W * (C + 2 + 1 [+ C with the warped stack]).
// Roofline: HBM (gather with high locality).
#include "common.cuh"

namespace cf {
namespace fwl {
constexpr int THREADS = 256;

struct alignas(16) Moments { double sum, sumsq; };

__global__ void __launch_bounds__(THREADS)
voxel_flow_warp_kernel(const float *__restrict__ voxel, const float *__restrict__ disp, int C, int H, int W, int reverse,
                       float *__restrict__ warped, float *__restrict__ summed, Moments *__restrict__ partial) {
    const int b = blockIdx.y;
    const int64_t plane = (int64_t)H * W;
    const int p = blockIdx.x * THREADS + threadIdx.x;
    float acc = 0.f;
    if (p < plane) {
        const int y = p / W, x = p - y * W;
        float dx = __ldg(disp + (int64_t)b * 2 * plane + p), dy = __ldg(disp + ((int64_t)b * 2 + 1) * plane + p);
        if (reverse) { dx = -dx; dy = -dy; }
        const double inc = 1.0 / ((double)C - 1.0);   // python float arithmetic of loss.py:46
        const float *vb = voxel + (int64_t)b * C * plane;
        for (int i = 0; i < C; ++i) {
            const float r = (float)(reverse ? 1.0 - (double)i * inc : (double)i * inc);
            // loss.py:53-60, then ATen grid_sampler_unnormalize(align_corners=True): ((g + 1) / 2) * (size - 1)
            const float wx = __fadd_rn((float)x, __fmul_rn(dx, r)), wy = __fadd_rn((float)y, __fmul_rn(dy, r));
            const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, wx), (float)W), 1.0f);
            const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, wy), (float)H), 1.0f);
            const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(W - 1));
            const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(H - 1));
            const float fx0 = floorf(ix), fy0 = floorf(iy);
            // float -> int of a wild coordinate is undefined: clamp first (anything that far out is all padding)
            const int x0 = (int)fminf(fmaxf(fx0, -2.f), (float)W), y0 = (int)fminf(fmaxf(fy0, -2.f), (float)H);
            const float ax1 = (fx0 + 1.f) - ix, ax0 = ix - fx0, ay1 = (fy0 + 1.f) - iy, ay0 = iy - fy0;
            const float *s = vb + (int64_t)i * plane;
            const bool xin0 = x0 >= 0 && x0 < W, xin1 = x0 + 1 >= 0 && x0 + 1 < W;
            const bool yin0 = y0 >= 0 && y0 < H, yin1 = y0 + 1 >= 0 && y0 + 1 < H;
            float v = 0.f;   // ATen order: nw, ne, sw, se, out-of-bounds taps skipped (zeros padding)
            if (xin0 && yin0) v += __ldg(s + (int64_t)y0 * W + x0) * (ax1 * ay1);
            if (xin1 && yin0) v += __ldg(s + (int64_t)y0 * W + x0 + 1) * (ax0 * ay1);
            if (xin0 && yin1) v += __ldg(s + (int64_t)(y0 + 1) * W + x0) * (ax1 * ay0);
            if (xin1 && yin1) v += __ldg(s + (int64_t)(y0 + 1) * W + x0 + 1) * (ax0 * ay0);
            if (warped) st_cs(warped + ((int64_t)b * C + i) * plane + p, v);
            acc += v;     // loss.py:66: channels added in order
        }
        summed[(int64_t)b * plane + p] = acc;
    }
    if (partial) {  // block moments of the summed image (fp64), fixed-shape tree
        double s = p < plane ? (double)acc : 0.0, q = s * s;
        s = warp_sum(s); q = warp_sum(q);
        __shared__ Moments sh[THREADS / 32];
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) sh[wid] = Moments{s, q};
        __syncthreads();
        if (threadIdx.x == 0) {
            Moments t = sh[0];
            for (int k = 1; k < THREADS / 32; ++k) { t.sum += sh[k].sum; t.sumsq += sh[k].sumsq; }
            partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
        }
    }
}

// mean and unbiased variance of n values from `count` block moments: single CTA, fixed order
__global__ void __launch_bounds__(THREADS) moments_finish_kernel(const Moments *__restrict__ partial, int count, double n,
                                                                 double *__restrict__ out) {
    double s = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < count; i += THREADS) { s += partial[i].sum; q += partial[i].sumsq; }
    s = warp_sum(s); q = warp_sum(q);
    __shared__ Moments sh[THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = Moments{s, q};
    __syncthreads();
    if (threadIdx.x == 0) {
        Moments t = sh[0];
        for (int k = 1; k < THREADS / 32; ++k) { t.sum += sh[k].sum; t.sumsq += sh[k].sumsq; }
        const double mean = t.sum / n;
        out[0] = mean;
        out[1] = n > 1.0 ? fmax(t.sumsq - t.sum * mean, 0.0) / (n - 1.0) : nan("");  // tensor.var(): correction = 1
    }
}
}  // namespace fwl
}  // namespace cf

extern "C" size_t cf_voxel_flow_warp_workspace_bytes(int B, int H, int W) {
    const size_t blocks = (size_t)cf::ceil_div((int64_t)H * W, cf::fwl::THREADS) * (size_t)(B > 0 ? B : 1);
    return blocks * sizeof(cf::fwl::Moments);
}

extern "C" int cf_voxel_flow_warp(const float *voxel, const float *displacement, int B, int C, int H, int W,
                                  int reverse_time, float *warped, float *summed, double *mean_var, void *ws,
                                  size_t ws_bytes, cf_stream_t stream_) {
    using namespace cf;
    if (int rc = check_device()) return rc;
    CF_REQUIRE(voxel && displacement && summed, CF_ERR_NULL, "cf_voxel_flow_warp: null pointer");
    CF_REQUIRE(B >= 0 && C >= 2 && H > 0 && W > 0 && B <= 65535, CF_ERR_INVALID_ARG,
               "cf_voxel_flow_warp: need B >= 0, C >= 2 (the reference divides by C - 1), H, W > 0; got B=%d C=%d H=%d W=%d", B, C, H, W);
    CF_REQUIRE((int64_t)H * W < (1ll << 31) - fwl::THREADS, CF_ERR_INVALID_ARG, "cf_voxel_flow_warp: plane too large");
    if (B == 0) return CF_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const unsigned bx = (unsigned)ceil_div((int64_t)H * W, fwl::THREADS);
    fwl::Moments *partial = nullptr;
    if (mean_var) {
        const size_t need = cf_voxel_flow_warp_workspace_bytes(B, H, W);
        CF_REQUIRE(ws && ws_bytes >= need, CF_ERR_WORKSPACE, "cf_voxel_flow_warp: workspace too small (%zu < %zu)", ws_bytes, need);
        CF_REQUIRE(aligned16(ws), CF_ERR_ALIGN, "cf_voxel_flow_warp: workspace not 16-byte aligned");
        partial = reinterpret_cast<fwl::Moments *>(ws);
    }
    fwl::voxel_flow_warp_kernel<<<dim3(bx, (unsigned)B), fwl::THREADS, 0, stream>>>(voxel, displacement, C, H, W, reverse_time != 0,
                                                                                    warped, summed, partial);
    CF_LAUNCH_CHECK("voxel_flow_warp_kernel");
    if (mean_var) {
        fwl::moments_finish_kernel<<<1, fwl::THREADS, 0, stream>>>(partial, (int)(bx * (unsigned)B), (double)B * H * W, mean_var);
        CF_LAUNCH_CHECK("moments_finish_kernel");
    }
    return CF_OK;
}
