// Standalone single-tile probe for the tcgen05 kind::tf32 MN-major path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/tc_probe scripts/tc_probe.cu
// One CTA, M=128, N=128, K=32.  Variants: operand fill (manual swizzled / TMA),
// LBO/SBO values, swizzle mode.  Prints the error of each variant and a few raw values.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 128, N = 128, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Variant {
    int fill;        // 0 manual, 1 TMA
    uint32_t lbo, sbo;
    int layout;      // descriptor layout type (2 = SW128, 0 = none)
    int swizzle;     // manual fill: apply 128B swizzle (1) or not (0)
    int kstep;       // bytes to advance the start address per K=8 MMA
    int a_major, b_major;  // 1 = MN-major (what we want)
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA_, const __grid_constant__ CUtensorMap tmB_,
             const __grid_constant__ CUtensorMap tmA32, const __grid_constant__ CUtensorMap tmB32,
             const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ C,
             float *__restrict__ smem_dump, Variant v) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sa = smem, *sb = smem + 16384;
    uint64_t *bar_tma = (uint64_t *)(smem + 32768), *bar_mma = bar_tma + 1;
    uint32_t *slot = (uint32_t *)(bar_tma + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    const CUtensorMap &tmA = v.swizzle == 2 ? tmA32 : tmA_;
    const CUtensorMap &tmB = v.swizzle == 2 ? tmB32 : tmB_;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_tma)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_mma)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;

    if (v.fill == 0) {
        // element (k, m) of A (MN-major: A[k][m]); atoms of 32 columns, 128-byte rows
        for (int idx = tid; idx < K * M; idx += 128) {
            const int k = idx / M, m = idx % M;
            const int atom = m / 32, c = m % 32;
            int off;
            if (v.swizzle == 2) {            // 128B swizzle with 32-byte atoms: Swizzle<2,5,2>
                const int chunk = (c / 8) ^ (k % 4);
                off = atom * 4096 + k * 128 + chunk * 32 + (c % 8) * 4;
            } else {
                int chunk = c / 4;
                if (v.swizzle) chunk ^= (k % 8);
                off = atom * 4096 + k * 128 + chunk * 16 + (c % 4) * 4;
            }
            *(float *)(sa + off) = A[k * M + m];
            *(float *)(sb + off) = B[k * N + m];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    } else {
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar_tma)), "r"(32768) : "memory");
            for (int a = 0; a < 4; ++a) {
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(smem_u32(sa + a * 4096)), "l"((uint64_t)&tmA), "r"(smem_u32(bar_tma)), "r"(32 * a), "r"(0) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(smem_u32(sb + a * 4096)), "l"((uint64_t)&tmB), "r"(smem_u32(bar_tma)), "r"(32 * a), "r"(0) : "memory");
            }
        }
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(bar_tma)) : "memory");
        }
        __syncthreads();
    }
    if (smem_dump) {
        for (int idx = tid; idx < 8192; idx += 128) smem_dump[idx] = ((float *)smem)[idx];
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)v.a_major << 15) | ((uint32_t)v.b_major << 16) |
                               ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int kk = 0; kk < K / 8; ++kk) {
            auto desc = [&](uint32_t addr) {
                uint64_t d = 0;
                d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
                d |= (uint64_t)(v.lbo >> 4) << 16;
                d |= (uint64_t)(v.sbo >> 4) << 32;
                d |= (uint64_t)1 << 46;
                d |= (uint64_t)v.layout << 61;
                return d;
            };
            const uint64_t da = desc(smem_u32(sa) + kk * v.kstep), db = desc(smem_u32(sb) + kk * v.kstep);
            const uint32_t acc = kk != 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar_mma)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar_mma)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = tid;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 32; ++q) C[row * N + c0 + q] = __uint_as_float(r[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    std::vector<float> hA(K * M), hB(K * N), ref(M * N), hC(M * N), dump(8192);
    srand(1);
    for (auto &x : hA) x = (float)(rand() % 17 - 8) * 0.25f;   // exactly representable in tf32
    for (auto &x : hB) x = (float)(rand() % 13 - 6) * 0.5f;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)hA[k * M + m] * hB[k * N + n];
            ref[m * N + n] = (float)s;
        }
    float *dA, *dB, *dC, *dDump;
    CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dC, hC.size() * 4));
    CK(cudaMalloc(&dDump, dump.size() * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));

    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fp;
    CUtensorMap tmA, tmB;
    cuuint64_t dims[2] = {(cuuint64_t)M, (cuuint64_t)K};
    cuuint64_t strides[1] = {(cuuint64_t)M * 4};
    cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
    CUresult r1 = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUtensorMap tmA32, tmB32;
    CUresult r3 = enc(&tmA32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r4 = enc(&tmB32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d %d %d %d\n", (int)r1, (int)r2, (int)r3, (int)r4);
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960));

    Variant vs[] = {
        // fill, lbo,  sbo,  layout, swz, kstep, amaj, bmaj
        {0, 4096, 512, 1, 2, 1024, 1, 1},    // SW128 with 32B atoms (layout type 1), manual fill
        {1, 4096, 512, 1, 2, 1024, 1, 1},    // same, TMA fill (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
        {0, 512, 4096, 1, 2, 1024, 1, 1},    // LBO/SBO swapped
        {1, 4096, 1024, 1, 2, 1024, 1, 1},   // SBO = 1024
        {0, 4096, 1024, 2, 1, 1024, 1, 1},   // first attempt: plain SW128 (gives zeros)
        {1, 4096, 1024, 2, 1, 1024, 1, 1},   // intended configuration, TMA fill
        {0, 1024, 4096, 2, 1, 1024, 1, 1},   // LBO/SBO swapped
        {1, 1024, 4096, 2, 1, 1024, 1, 1},
        {0, 4096, 1024, 2, 1, 1024, 0, 0},   // interpret as K-major (expected wrong)
        {0, 4096, 128, 0, 0, 1024, 1, 1},    // no swizzle: rows of 128B, "interleave" style
        {0, 128, 4096, 0, 0, 1024, 1, 1},
    };
    for (auto &v : vs) {
        CK(cudaMemset(dC, 0xff, hC.size() * 4));
        probe_kernel<<<1, 128, 40960>>>(tmA, tmB, tmA32, tmB32, dA, dB, dC, dDump, v);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant fill=%d lbo=%u sbo=%u: CUDA error %s\n", v.fill, v.lbo, v.sbo, cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(dump.data(), dDump, dump.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0, maxref = 0;
        int nz = 0, nan = 0;
        for (int i = 0; i < M * N; ++i) {
            if (isnan(hC[i])) { nan++; continue; }
            maxerr = fmax(maxerr, fabs((double)hC[i] - ref[i]));
            maxref = fmax(maxref, fabs((double)ref[i]));
            nz += hC[i] != 0.f;
        }
        // check smem content against the intended layout
        int smem_bad = 0;
        for (int k = 0; k < K; ++k)
            for (int m = 0; m < M; ++m) {
                const int atom = m / 32, c = m % 32;
                int off;
                if (v.swizzle == 2) off = (atom * 4096 + k * 128 + (((c / 8) ^ (k % 4)) * 32) + (c % 8) * 4) / 4;
                else off = (atom * 4096 + k * 128 + (((c / 4) ^ (k % 8)) * 16) + (c % 4) * 4) / 4;
                if (dump[off] != hA[k * M + m]) smem_bad++;
            }
        printf("fill=%d lbo=%u sbo=%u layout=%d swz=%d major=%d%d : maxerr=%.4g maxref=%.4g nonzero=%d nan=%d smem_mismatch=%d  C[0][0..3]=%g %g %g %g ref=%g %g %g %g\n",
               v.fill, v.lbo, v.sbo, v.layout, v.swizzle, v.a_major, v.b_major, maxerr, maxref, nz, nan, smem_bad,
               hC[0], hC[1], hC[2], hC[3], ref[0], ref[1], ref[2], ref[3]);
    }
    return 0;
}
