// Correlation pyramid lookup (part 3b of the hot path).
//
// Replaces CorrBlock.__call__ (ERAFT/corr.py:29-50 == DCEIFlow/core/corr/
// raft_corr.py:32-54) and its bilinear_sampler (ERAFT/utils.py:7-21): for each
// query pixel q and pyramid level l the reference samples a (2r+1)^2 window
// around coords[q]/2^l with grid_sample(align_corners=True, zeros padding).
// The window offsets are integers, so all (2r+1)^2 samples of a level share
// one pair of fractional weights and read one (2r+2)^2 patch: the kernel loads
// the patch once (zero-filled outside the map) and forms the samples from
// shared memory.  The window is TRANSPOSED as in the reference:
// channel l*(2r+1)^2 + i*(2r+1) + j samples (x + i - r, y + j - r).
//
// Data layout: pyramid[l] [B*N, 1, h>>l, w>>l], coords [B,2,h,w], out
// [B, L*(2r+1)^2, h, w]; fp32.  Roofline: HBM (L2 when the pyramid was just
// built).  Algorithmic bytes per query: 4*L*(2r+1)^2 written + 8 read (coords)
// + 4*L*(2r+2)^2 read (patches) = 2904 B at L=4, r=4.
//
// Mapping: one CTA = 32 consecutive queries of one batch item; one warp walks
// 4 queries, its lanes fetch the L patches of a query with all loads in flight,
// then produce the L*(2r+1)^2 samples.  Results are staged in a padded shared
// tile [channel][query] so that the global store is channel-major and fully
// coalesced (32 consecutive queries = one 128-byte line per channel); writing
// them straight from the lanes would scatter 4-byte stores N*4 bytes apart.
#include "common.cuh"

namespace cf {

struct Pyramid {
    const float *ptr[CF_CORR_MAX_LEVELS];
    int H[CF_CORR_MAX_LEVELS];
    int W[CF_CORR_MAX_LEVELS];
};

constexpr int kQueriesPerCta = 32;
constexpr int kLookupWarps = 8;

// RADIUS > 0: compile-time radius; RADIUS == 0: use the runtime argument.
template <int RADIUS>
__global__ void __launch_bounds__(kLookupWarps * 32)
corr_lookup_kernel(Pyramid pyr, const float *__restrict__ coords, float *__restrict__ out,
                   int N, int levels, int radius_rt) {
    extern __shared__ float smem[];
    const int r = RADIUS > 0 ? RADIUS : radius_rt;
    const int K = 2 * r + 1, P = K + 1, KK = K * K, PP = P * P;
    const int C = levels * KK;
    constexpr int TS = kQueriesPerCta + 1;  // padded row: conflict-free column writes
    float *tile = smem;                     // [C][TS]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *patch = smem + (size_t)C * TS + (size_t)warp * levels * PP;  // [levels][P][P]

    const int b = blockIdx.y;
    const int q0 = blockIdx.x * kQueriesPerCta;
    const float *cb = coords + (size_t)b * 2 * N;

    for (int qi = warp; qi < kQueriesPerCta; qi += kLookupWarps) {
        const int q = q0 + qi;
        if (q >= N) break;
        // clamp keeps (int) conversions defined for wild coordinates; anything
        // this far out samples only zero padding anyway
        const float cx = fminf(fmaxf(__ldg(cb + q), -1.0e6f), 1.0e6f);
        const float cy = fminf(fmaxf(__ldg(cb + N + q), -1.0e6f), 1.0e6f);
        const size_t map = (size_t)b * N + q;

        for (int e = lane; e < levels * PP; e += 32) {
            const int l = e / PP, rem = e - l * PP;
            const int py = rem / P, px = rem - py * P;
            const float inv = 1.f / (float)(1 << l);  // exact: coords / 2**l
            const int X = (int)floorf(cx * inv) - r + px;
            const int Y = (int)floorf(cy * inv) - r + py;
            const int Hl = pyr.H[l], Wl = pyr.W[l];
            float v = 0.f;
            if (X >= 0 && X < Wl && Y >= 0 && Y < Hl)
                v = __ldg(pyr.ptr[l] + (map * Hl + Y) * Wl + X);
            patch[e] = v;
        }
        __syncwarp();
        for (int o = lane; o < C; o += 32) {
            const int l = o / KK, rem = o - l * KK;
            const int i = rem / K, j = rem - i * K;  // i -> x offset, j -> y offset
            const float inv = 1.f / (float)(1 << l);
            const float sx = cx * inv, sy = cy * inv;
            const float fx = sx - floorf(sx), fy = sy - floorf(sy);
            const float *pp = patch + l * PP + j * P + i;
            const float v00 = pp[0], v01 = pp[1], v10 = pp[P], v11 = pp[P + 1];
            float acc = v00 * ((1.f - fx) * (1.f - fy));
            acc += v01 * (fx * (1.f - fy));
            acc += v10 * ((1.f - fx) * fy);
            acc += v11 * (fx * fy);
            tile[o * TS + qi] = acc;
        }
        __syncwarp();
    }
    __syncthreads();
    const int q = q0 + lane;
    if (q < N) {
        float *ob = out + (size_t)b * C * N + q;
        for (int ch = warp; ch < C; ch += kLookupWarps) ob[(size_t)ch * N] = tile[ch * TS + lane];
    }
}

}  // namespace cf

extern "C" int cf_corr_lookup(const float *const *pyramid, const float *coords, int B, int h, int w,
                              int levels, int radius, float *out, cf_stream_t stream_) {
    using namespace cf;
    if (int rc = check_device()) return rc;
    CF_REQUIRE(pyramid && coords && out, CF_ERR_NULL, "cf_corr_lookup: null pointer");
    CF_REQUIRE(levels >= 1 && levels <= CF_CORR_MAX_LEVELS, CF_ERR_INVALID_ARG,
               "cf_corr_lookup: levels=%d not in [1,%d]", levels, CF_CORR_MAX_LEVELS);
    CF_REQUIRE(radius >= 0 && radius <= 8, CF_ERR_INVALID_ARG, "cf_corr_lookup: radius=%d not in [0,8]", radius);
    CF_REQUIRE(B >= 0 && B <= 65535 && h > 0 && w > 0, CF_ERR_INVALID_ARG, "cf_corr_lookup: bad shape B=%d h=%d w=%d", B, h, w);
    CF_REQUIRE((h >> (levels - 1)) >= 1 && (w >> (levels - 1)) >= 1, CF_ERR_INVALID_ARG,
               "cf_corr_lookup: %dx%d feature map has no level %d", h, w, levels - 1);
    if (B == 0) return CF_OK;
    Pyramid pyr{};
    for (int l = 0; l < levels; ++l) {
        CF_REQUIRE(pyramid[l], CF_ERR_NULL, "cf_corr_lookup: pyramid[%d] is null", l);
        pyr.ptr[l] = pyramid[l];
        pyr.H[l] = h >> l;
        pyr.W[l] = w >> l;
    }
    const int N = h * w, K = 2 * radius + 1, P = K + 1;
    const int C = levels * K * K;
    const size_t smem = ((size_t)C * (kQueriesPerCta + 1) + (size_t)kLookupWarps * levels * P * P) * sizeof(float);
    CF_REQUIRE(smem <= 200 * 1024, CF_ERR_INVALID_ARG, "cf_corr_lookup: levels*radius too large for shared memory");
    cudaStream_t stream = (cudaStream_t)stream_;
    dim3 grid((unsigned)ceil_div(N, kQueriesPerCta), B);
    int dev = 0;
    CF_CUDA(cudaGetDevice(&dev));
    static bool smem_opt_in[2][64] = {};
    auto kern = radius == 4 ? corr_lookup_kernel<4> : corr_lookup_kernel<0>;
    bool &done = smem_opt_in[radius == 4][dev & 63];
    if (!done) {
        CF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        done = true;
    }
    kern<<<grid, kLookupWarps * 32, smem, stream>>>(pyr, coords, out, N, levels, radius);
    CF_LAUNCH_CHECK("corr_lookup_kernel");
    return CF_OK;
}
