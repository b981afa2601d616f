// Packed event ingest: 8 bytes per event instead of the reference's
// 32-byte fp64 row (data_readers/event_readers.py:6-47 delivers [N,4] float64 (t, x, y, p)).
//
//   word 0  float32  t_rel = (float)(t - t_first_of_window): the fp64 subtraction happens BEFORE the
//                    rounding, so absolute stamps (~1e9 us) cost no precision
//   word 1  uint32   x | y << 16 | p << 31          (x < 65536, y < 32768, p = 1 for positive polarity)
//
// cf_events_pack       device-side packer: fp64 rows -> packed (one pass, 32 B read + 8 B written per event)
// cf_voxel_bin_packed  ATOMIC-mode voxel grid (+ fused event_preprocess) from packed events: the kernel
//                      reads 8 B per event (4x fewer event bytes than cf_voxel_bin) and does the time
//                      normalisation t* = (nb-1) * t_rel / dT in fp32 (mul, then div, like
//                      utils/event_process.py:46-49).  Against the fp64 reference the bin position moves
//                      by <= ~4e-7 and, the temporal bilinear weights being continuous across bin
//                      boundaries, every cell stays within the atomic-mode tolerance 1e-5 * (sum|w| + 1).
//                      There is no deterministic (bit-exact) mode for packed input: the fp64 stamps
//                      that define the reference's bins are gone.
#include "voxel_common.cuh"

namespace cf {

// from voxel.cu: zero -> [scatter] -> statistics -> normalise-in-place, shared with the fp64 path
int run_preprocess_shared(const float *in, float *out, int B, int64_t cells, int preprocess, float hot_thr, void *ws,
                          size_t ws_bytes, cudaStream_t stream);

namespace vp {
constexpr int THREADS = 256;
constexpr int UNROLL = 4;

__device__ __forceinline__ int find_window(int64_t i, const int64_t *__restrict__ off, int B) {
    int lo = 0, hi = B - 1;  // last b with off[b] <= i
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(off + mid) <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(THREADS)
events_pack_kernel(const double *__restrict__ ev, const int64_t *__restrict__ off, int64_t total, int B,
                   uint2 *__restrict__ packed) {
    int b = -1;
    int64_t begin = 0, end = 0;
    double t0 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * THREADS) {
        const Event e = load_event(ev, i);
        if (b < 0 || i < begin || i >= end) {
            b = find_window(i, off, B);
            begin = __ldg(off + b);
            end = __ldg(off + b + 1);
            t0 = __ldg(ev + 4 * begin);
        }
        // out-of-range coordinates are kept representable and marked invalid (x = 0xffff): the voxel kernel
        // drops them like cf_voxel_bin drops out-of-grid events
        const bool ok = e.x >= 0.0 && e.x < 65535.0 && e.y >= 0.0 && e.y < 32768.0;
        const uint32_t x = ok ? (uint32_t)e.x : 0xffffu, y = ok ? (uint32_t)e.y : 0u;
        const uint32_t p = e.p > 0.0 ? 1u : 0u;
        packed[i] = make_uint2(__float_as_uint((float)__dsub_rn(e.t, t0)), x | (y << 16) | (p << 31));
    }
}

__global__ void __launch_bounds__(THREADS)
voxel_scatter_packed_kernel(const uint2 *__restrict__ packed, const int64_t *__restrict__ off, int B, int nb, int H, int W,
                            float *__restrict__ out) {
    const int64_t ev_end = __ldg(off + B);
    const int64_t plane = (int64_t)H * W;
    constexpr int64_t kTile = THREADS * UNROLL;
    int b = -1;
    int64_t begin = 0, end = 0;
    float span = 1.f;
    for (int64_t tile = (int64_t)blockIdx.x * kTile; tile < ev_end; tile += (int64_t)gridDim.x * kTile) {
        uint2 e[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {  // all loads in flight first
            const int64_t i = tile + k * THREADS + threadIdx.x;
            if (i < ev_end) e[k] = __ldg(packed + i);
        }
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            const int64_t i = tile + k * THREADS + threadIdx.x;
            if (i >= ev_end) break;
            if (b < 0 || i < begin || i >= end) {
                b = find_window(i, off, B);
                begin = __ldg(off + b);
                end = __ldg(off + b + 1);
                span = __uint_as_float(__ldg(packed + end - 1).x);   // t_rel of the window's last event
                if (span == 0.f) span = 1.f;                          // event_process.py:43-44
            }
            const float tn = __fdiv_rn(__fmul_rn((float)(nb - 1), __uint_as_float(e[k].x)), span);
            const float lo = floorf(tn);
            const int x = (int)(e[k].y & 0xffffu), y = (int)((e[k].y >> 16) & 0x7fffu);
            if (!(lo >= 0.f && lo < (float)nb) || x >= W || y >= H) continue;
            const int bin = (int)lo;
            const float f = __fsub_rn(tn, lo);
            const float s = (e[k].y >> 31) ? 1.f : -1.f;             // event_process.py:51 (p == 0 -> -1)
            float *cell = out + ((int64_t)b * nb + bin) * plane + (int64_t)y * W + x;
            atomicAdd(cell, __fmul_rn(s, __fsub_rn(1.f, f)));
            if (bin + 1 < nb) atomicAdd(cell + plane, __fmul_rn(s, f));
        }
    }
}
}  // namespace vp
}  // namespace cf

extern "C" int cf_events_pack(const double *events, const int64_t *offsets, int64_t total, int B, void *packed,
                              cf_stream_t stream_) {
    using namespace cf;
    if (int rc = check_device()) return rc;
    CF_REQUIRE(offsets && packed, CF_ERR_NULL, "cf_events_pack: null pointer");
    CF_REQUIRE(total == 0 || events, CF_ERR_NULL, "cf_events_pack: events is null");
    CF_REQUIRE(B >= 1 && total >= 0, CF_ERR_INVALID_ARG, "cf_events_pack: bad sizes B=%d total=%lld", B, (long long)total);
    CF_REQUIRE(total == 0 || aligned16(events), CF_ERR_ALIGN, "cf_events_pack: events not 16-byte aligned");
    CF_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 7u) == 0, CF_ERR_ALIGN, "cf_events_pack: packed not 8-byte aligned");
    if (total == 0) return CF_OK;
    int64_t blocks = ceil_div(total, vp::THREADS);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    vp::events_pack_kernel<<<(unsigned)blocks, vp::THREADS, 0, (cudaStream_t)stream_>>>(events, offsets, total, B,
                                                                                        reinterpret_cast<uint2 *>(packed));
    CF_LAUNCH_CHECK("events_pack_kernel");
    return CF_OK;
}

extern "C" int cf_voxel_bin_packed(const void *packed, const int64_t *offsets, int64_t total, int B, int nb, int H, int W,
                                   int preprocess, float hot_thr, float *out, void *ws, size_t ws_bytes,
                                   cf_stream_t stream_) {
    using namespace cf;
    if (int rc = check_device()) return rc;
    CF_REQUIRE(out && offsets, CF_ERR_NULL, "cf_voxel_bin_packed: null pointer");
    CF_REQUIRE(total == 0 || packed, CF_ERR_NULL, "cf_voxel_bin_packed: packed is null");
    CF_REQUIRE(nb > 0 && H > 0 && W > 0 && H <= 32768 && W <= 65535, CF_ERR_INVALID_ARG,
               "cf_voxel_bin_packed: num_bins > 0, 0 < width <= 65535, 0 < height <= 32768 required");
    CF_REQUIRE(B >= 0 && total >= 0, CF_ERR_INVALID_ARG, "cf_voxel_bin_packed: negative size");
    CF_REQUIRE(preprocess >= CF_PRE_NONE && preprocess <= CF_PRE_MAXMIN, CF_ERR_INVALID_ARG, "cf_voxel_bin_packed: bad preprocess %d", preprocess);
    CF_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 7u) == 0, CF_ERR_ALIGN, "cf_voxel_bin_packed: packed not 8-byte aligned");
    if (B == 0) return CF_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t cells = (int64_t)nb * H * W;
    CF_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * cells, stream));
    if (total > 0) {
        int64_t blocks = ceil_div(total, vp::THREADS * vp::UNROLL);
        const int64_t cap = (int64_t)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        vp::voxel_scatter_packed_kernel<<<(unsigned)blocks, vp::THREADS, 0, stream>>>(
            reinterpret_cast<const uint2 *>(packed), offsets, B, nb, H, W, out);
        CF_LAUNCH_CHECK("voxel_scatter_packed_kernel");
    }
    if (preprocess != CF_PRE_NONE) return run_preprocess_shared(out, out, B, cells, preprocess, hot_thr, ws, ws_bytes, stream);
    return CF_OK;
}
