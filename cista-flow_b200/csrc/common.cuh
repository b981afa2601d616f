// Shared host/device helpers for libcistaflow.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "cistaflow.h"

namespace cf {

// ---- host-side error plumbing ------------------------------------------------
void set_error(const char *fmt, ...);
int check_device();  // CF_OK or CF_ERR_ARCH/CF_ERR_CUDA (cached per device)
int sm_count();      // SM count of the current device (148 on B200)
void count_launch(const char *kernel); // bumps cf_launch_count(), remembers the name for cf_last_kernel()

#define CF_REQUIRE(cond, code, ...)      \
    do {                                 \
        if (!(cond)) {                   \
            cf::set_error(__VA_ARGS__);  \
            return (code);               \
        }                                \
    } while (0)

#define CF_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (expr);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            cf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),      \
                          __FILE__, __LINE__);                                          \
            return CF_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

#define CF_LAUNCH_CHECK(name)                                                           \
    do {                                                                                \
        cf::count_launch(name);                                                             \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) {                                                       \
            cf::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));    \
            return CF_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device helpers -----------------------------------------------------------
__device__ __forceinline__ float ld_nc(const float *p) { return __ldg(p); }

// streaming (evict-first) stores for outputs that are written once
__device__ __forceinline__ void st_cs(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs4(float4 *p, float4 v) { __stcs(p, v); }

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

#ifdef CF_TRACE
// experiment builds only (build.build_variant('trace', ['CF_TRACE'])): per-CTA timelines in globaltimer ns,
// read back by scripts/*_trace.py through cf_trace_buffer()
// (one copy per translation unit -- no relocatable device code; each traced .cu exports its own setter)
static __device__ unsigned long long *g_trace = nullptr;
constexpr int kTraceSlots = 256;
#define CF_DEFINE_TRACE_SETTER(name)                                                              \
    extern "C" __attribute__((visibility("default"))) int name(unsigned long long *buf) {        \
        return (int)cudaMemcpyToSymbol(cf::g_trace, &buf, sizeof(buf));                           \
    }
__device__ __forceinline__ void trace_at(int slot) {
    if (g_trace && slot < kTraceSlots) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kTraceSlots + slot] = t;
    }
}
#define CF_TRACE_AT(slot) cf::trace_at(slot)
#else
#define CF_TRACE_AT(slot) ((void)0)
#define CF_DEFINE_TRACE_SETTER(name)
#endif

}  // namespace cf
