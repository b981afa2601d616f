// Round-2 probe: what the B200 memory system gives the voxel path.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/atomics_probe scripts/experiments/atomics_probe.cu
// Measures (CUDA events, best of 5 after warm-up):
//   1. fp32 RED throughput on spread addresses: scalar pairs a plane apart (what voxel_scatter issues today),
//      one v2 per event, one v4 per event -- grid regions of 49 MB (8 windows of 5x480x640) and 393 MB (64)
//   2. the same with returning atomics (ATOMG)
//   3. pass bandwidth over a region that is L2-resident (49 MB) vs not (1 GB): write, read, read+write
//   4. shared-memory fp32 atomicAdd (ATOMS.CAST.SPIN) and integer ATOMS throughput per SM
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void red_scalar_pairs(float *g, const uint32_t *idx, const float *w, int64_t n, int64_t plane) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t c = idx[i];
        const float v = w[i];
        atomicAdd(g + c, v);
        atomicAdd(g + c + plane, 1.f - v);
    }
}
__global__ void red_scalar_one(float *g, const uint32_t *idx, const float *w, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(g + idx[i], w[i]);
}
__global__ void red_v2(float *g, const uint32_t *idx, const float *w, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = w[i];
        float *p = g + ((size_t)idx[i] & ~(size_t)1);
        asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v), "f"(1.f - v) : "memory");
    }
}
__global__ void red_v4(float *g, const uint32_t *idx, const float *w, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = w[i];
        float *p = g + ((size_t)idx[i] & ~(size_t)3);
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v), "f"(1.f - v), "f"(0.f), "f"(0.f) : "memory");
    }
}
__global__ void atom_scalar_pairs(float *g, const uint32_t *idx, const float *w, int64_t n, int64_t plane, float *sink) {
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t c = idx[i];
        const float v = w[i];
        acc += atomicAdd(g + c, v);
        acc += atomicAdd(g + c + plane, 1.f - v);
    }
    if (acc == 123.456f) *sink = acc;
}
__global__ void pass_write(float4 *g, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
        g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
__global__ void pass_read(const float4 *g, int64_t n4, float *sink) {
    float acc = 0.f;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    for (; i + 3 * st < n4; i += 4 * st) {
        const float4 a = g[i], b = g[i + st], c = g[i + 2 * st], d = g[i + 3 * st];
        acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.x + d.y + d.z + d.w;
    }
    for (; i < n4; i += st) { const float4 a = g[i]; acc += a.x + a.y + a.z + a.w; }
    if (acc == 123.456f) *sink = acc;
}
__global__ void pass_rw(float4 *g, int64_t n4, float s) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    for (; i + 3 * st < n4; i += 4 * st) {
        float4 a = g[i], b = g[i + st], c = g[i + 2 * st], d = g[i + 3 * st];
        a.x *= s; a.y *= s; a.z *= s; a.w *= s; b.x *= s; b.y *= s; b.z *= s; b.w *= s;
        c.x *= s; c.y *= s; c.z *= s; c.w *= s; d.x *= s; d.y *= s; d.z *= s; d.w *= s;
        g[i] = a; g[i + st] = b; g[i + 2 * st] = c; g[i + 3 * st] = d;
    }
    for (; i < n4; i += st) { float4 a = g[i]; a.x *= s; g[i] = a; }
}
__global__ void smem_atomics(const uint32_t *idx, const float *w, int per_cta, float *sink, int use_int) {
    extern __shared__ float s[];
    const int cells = 48 * 1024;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) s[i] = 0.f;
    __syncthreads();
    const uint32_t *id = idx + (size_t)blockIdx.x * per_cta;
    const float *ww = w + (size_t)blockIdx.x * per_cta;
    if (use_int) {
        for (int i = threadIdx.x; i < per_cta; i += blockDim.x) atomicAdd(reinterpret_cast<int *>(s) + id[i] % cells, 3);
    } else {
        for (int i = threadIdx.x; i < per_cta; i += blockDim.x) atomicAdd(s + id[i] % cells, ww[i]);
    }
    __syncthreads();
    if (s[threadIdx.x] == 123.456f) *sink = 1.f;
}

template <typename F>
static float time_best(F f, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    const int64_t plane = 480 * 640;
    float *sink;
    CK(cudaMalloc(&sink, 4));
    for (int windows : {8, 64}) {
        const int64_t cells = (int64_t)windows * 5 * plane;
        const int64_t n = (int64_t)windows * 100000;
        float *g;
        CK(cudaMalloc(&g, cells * 4 + 64));
        CK(cudaMemset(g, 0, cells * 4));
        std::vector<uint32_t> hidx(n);
        std::vector<float> hw(n);
        uint64_t s = 88172645463325252ull;
        for (int64_t i = 0; i < n; ++i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            const int64_t b = i / 100000;   // window-local like the real scatter: events of a window hit that window's grid
            hidx[i] = (uint32_t)(b * 5 * plane + (s % (uint64_t)(4 * plane)));
            hw[i] = (float)((s >> 40) & 1023) / 1024.f;
        }
        uint32_t *idx; float *w;
        CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&w, n * 4));
        CK(cudaMemcpy(idx, hidx.data(), n * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(w, hw.data(), n * 4, cudaMemcpyHostToDevice));
        for (int blocks : {148 * 2, 148 * 8}) {
            const int th = 256;
            float t;
            t = time_best([&] { red_scalar_pairs<<<blocks, th>>>(g, idx, w, n, plane); });
            printf("windows %2d grid %4d  RED scalar pair : %8.2f us  %7.1f G red/s  %6.1f G ev/s\n", windows, blocks, t * 1e3, 2 * n / t / 1e6, n / t / 1e6);
            t = time_best([&] { red_scalar_one<<<blocks, th>>>(g, idx, w, n); });
            printf("windows %2d grid %4d  RED scalar one  : %8.2f us  %7.1f G red/s\n", windows, blocks, t * 1e3, n / t / 1e6);
            t = time_best([&] { red_v2<<<blocks, th>>>(g, idx, w, n); });
            printf("windows %2d grid %4d  RED v2 per event: %8.2f us  %7.1f G red/s\n", windows, blocks, t * 1e3, n / t / 1e6);
            t = time_best([&] { red_v4<<<blocks, th>>>(g, idx, w, n); });
            printf("windows %2d grid %4d  RED v4 per event: %8.2f us  %7.1f G red/s\n", windows, blocks, t * 1e3, n / t / 1e6);
            t = time_best([&] { atom_scalar_pairs<<<blocks, th>>>(g, idx, w, n, plane, sink); });
            printf("windows %2d grid %4d  ATOM scalar pair: %8.2f us  %7.1f G atom/s\n", windows, blocks, t * 1e3, 2 * n / t / 1e6);
        }
        // memset + scatter back to back (what the product does), and the region passes
        float t = time_best([&] { CK(cudaMemsetAsync(g, 0, cells * 4)); });
        printf("windows %2d cudaMemsetAsync %lld MB: %8.2f us  %7.1f GB/s\n", windows, (long long)(cells * 4 >> 20), t * 1e3, cells * 4 / t / 1e6);
        for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
            t = time_best([&] { pass_write<<<blocks, 256>>>((float4 *)g, cells / 4); });
            printf("windows %2d grid %4d  write pass: %8.2f us  %7.1f GB/s\n", windows, blocks, t * 1e3, cells * 4 / t / 1e6);
            t = time_best([&] { pass_read<<<blocks, 256>>>((const float4 *)g, cells / 4, sink); });
            printf("windows %2d grid %4d  read pass : %8.2f us  %7.1f GB/s\n", windows, blocks, t * 1e3, cells * 4 / t / 1e6);
            t = time_best([&] { pass_rw<<<blocks, 256>>>((float4 *)g, cells / 4, 1.0001f); });
            printf("windows %2d grid %4d  r+w pass  : %8.2f us  %7.1f GB/s (read+write bytes)\n", windows, blocks, t * 1e3, 2 * cells * 4 / t / 1e6);
        }
        CK(cudaFree(g)); CK(cudaFree(idx)); CK(cudaFree(w));
    }
    {   // shared-memory atomics: 148 CTAs x 1024 threads, 200k adds per CTA into 48k cells (192 KB)
        const int per = 200000, ctas = 148;
        std::vector<uint32_t> hidx((size_t)per * ctas);
        std::vector<float> hw((size_t)per * ctas, 0.5f);
        uint64_t s = 1234567ull;
        for (auto &v : hidx) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (uint32_t)(s >> 16); }
        uint32_t *idx; float *w;
        CK(cudaMalloc(&idx, hidx.size() * 4)); CK(cudaMalloc(&w, hw.size() * 4));
        CK(cudaMemcpy(idx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(smem_atomics, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024));
        for (int use_int : {0, 1}) {
            float t = time_best([&] { smem_atomics<<<ctas, 1024, 192 * 1024>>>(idx, w, per, sink, use_int); });
            printf("smem %s atomics: %8.2f us for %d per CTA -> %6.2f adds/clk/SM-ish, %7.1f G/s chip\n", use_int ? "int" : "f32",
                   t * 1e3, per, per / (t * 1e-3 * 1.9e9), (double)per * ctas / t / 1e6);
        }
    }
    return 0;
}
