// Round-2 probe: how fast can ONE SM's TMA unit load operand stages of the shapes corr_build_tc.cu uses?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_probe scripts/experiments/tma_probe.cu -lcuda
// Every CTA (1 per SM, 148 CTAs) runs a producer thread that keeps a ring of S stages full (no consumer: the stage
// is re-armed as soon as it has landed), walking over a [rows = B*D][N] feature map like the GEMM's main loop does.
// Reported: GB/s per SM and chip-wide, for
//   tf32_3d   fp32 words, TFLOAT32 map {32 cols, 32 rows, 9 atoms}, SWIZZLE_128B_ATOM_32B   (the TF32 kernel's stage)
//   f32_3d    the same with a plain FLOAT32 map (no rounding in flight)
//   f32_3d_sw128   FLOAT32, plain SWIZZLE_128B
//   tf32_2d   nine 2-D boxes {32 cols, 32 rows} per stage
//   f16_3d    fp16, {64 cols, 32 rows, 5 atoms}, SWIZZLE_128B                                  (the fp16 kernel's stage)
//   f16_2d    five 2-D boxes {64 cols, 32 rows}
//   f16_3d_k64  fp16, {64 cols, 64 rows, 5 atoms} (twice the K rows per instruction)
//   f32_rows  fp32 {32 cols, 256 rows} (one atom, many rows: the A-resident slice shape)
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    return reinterpret_cast<EncodeTiledFn>(p);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *m, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *m, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

struct Cfg {
    int mode3d;        // 1: one 3-D instruction per stage, 0: `boxes` 2-D instructions
    int boxes;         // atoms per stage
    int box_cols;      // elements per box row
    int box_rows;
    int elt;           // bytes per element
    int stages;
    int iters;         // stages loaded per CTA
    int N;             // columns of the map
    int rows_total;    // rows of the map
};

__global__ void __launch_bounds__(64, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap tmap, const Cfg c, unsigned long long *cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[16];
    const int box_bytes = c.box_cols * c.box_rows * c.elt;
    const int stage_bytes = c.boxes * box_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < c.stages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        // walk like the GEMM: column block fixed per "tile", K rows advance by box_rows per stage
        const int col_blocks = c.N / (c.boxes * c.box_cols);
        int issued = 0, waited = 0;
        uint32_t phase_bits = 0;
        while (waited < c.iters) {
            while (issued < c.iters && issued - waited < c.stages) {
                const int s = issued % c.stages;
                const int tile = (blockIdx.x * 7 + issued / 8) % col_blocks;
                const int krow = ((blockIdx.x * 13 + issued) * c.box_rows) % (c.rows_total - c.box_rows);
                const int col0 = tile * c.boxes * c.box_cols;
                mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
                if (c.mode3d) {
                    tma_load_3d(smem + s * stage_bytes, &tmap, 0, krow, col0 / c.box_cols, &full[s]);
                } else {
                    for (int a = 0; a < c.boxes; ++a) tma_load_2d(smem + s * stage_bytes + a * box_bytes, &tmap, col0 + a * c.box_cols, krow, &full[s]);
                }
                ++issued;
            }
            const int s = waited % c.stages;
            while (!mbar_try_wait(&full[s], (phase_bits >> s) & 1u)) {}
            phase_bits ^= 1u << s;
            ++waited;
        }
        cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
}

static void run(const char *name, CUtensorMapDataType dt, CUtensorMapSwizzle sw, int mode3d, int boxes, int box_cols, int box_rows,
                int elt, int stages, void *base, int rows_total, int N) {
    EncodeTiledFn fn = encode_fn();
    CUtensorMap m;
    CUresult r;
    if (mode3d) {
        cuuint64_t dims[3] = {(cuuint64_t)box_cols, (cuuint64_t)rows_total, (cuuint64_t)(N / box_cols)};
        cuuint64_t strides[2] = {(cuuint64_t)N * elt, (cuuint64_t)box_cols * elt};
        cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, (cuuint32_t)boxes};
        cuuint32_t es[3] = {1, 1, 1};
        r = fn(&m, dt, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)rows_total};
        cuuint64_t strides[1] = {(cuuint64_t)N * elt};
        cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
        cuuint32_t es[2] = {1, 1};
        r = fn(&m, dt, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("%-14s encode failed (%d)\n", name, (int)r); return; }
    Cfg c{mode3d, boxes, box_cols, box_rows, elt, stages, 2000, N, rows_total};
    const int stage_bytes = boxes * box_cols * box_rows * elt;
    const int smem = stages * stage_bytes + 1024;
    if (smem > 226 * 1024) { printf("%-14s x%d stages: does not fit\n", name, stages); return; }
    CK(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    unsigned long long *cyc;
    CK(cudaMalloc(&cyc, 148 * 8));
    for (int ctas : {1, 148}) {
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        tma_probe_kernel<<<ctas, 64, smem>>>(m, c, cyc);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a));
        tma_probe_kernel<<<ctas, 64, smem>>>(m, c, cyc);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        const double bytes = (double)c.iters * stage_bytes;
        printf("%-14s stage %6d B x%2d stages, %3d CTAs: %7.1f GB/s per SM, %8.1f GB/s chip, %6.0f ns per stage\n", name, stage_bytes,
               stages, ctas, bytes / (ms * 1e-3) / 1e9, bytes * ctas / (ms * 1e-3) / 1e9, ms * 1e6 / c.iters);
    }
    CK(cudaFree(cyc));
}

int main() {
    const int B = 8, D = 256, N = 4800;          // 8 x 60x80 feature maps
    const int rows = B * D;
    float *f32;
    __half *f16;
    CK(cudaMalloc(&f32, (size_t)rows * N * 4));
    CK(cudaMalloc(&f16, (size_t)rows * N * 2));
    CK(cudaMemset(f32, 0, (size_t)rows * N * 4));
    CK(cudaMemset(f16, 0, (size_t)rows * N * 2));
    for (int stages : {2, 5, 8}) {
        if (stages <= 5) {
            run("tf32_3d", CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 1, 9, 32, 32, 4, stages, f32, rows, N);
            run("f32_3d", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 1, 9, 32, 32, 4, stages, f32, rows, N);
            run("f32_3d_sw128", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B, 1, 9, 32, 32, 4, stages, f32, rows, N);
            run("tf32_2d", CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 0, 9, 32, 32, 4, stages, f32, rows, N);
            run("f32_rows", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B, 0, 1, 32, 256, 4, stages, f32, rows, N);
        }
        run("f16_3d", CU_TENSOR_MAP_DATA_TYPE_FLOAT16, CU_TENSOR_MAP_SWIZZLE_128B, 1, 5, 64, 32, 2, stages, f16, rows, N);
        run("f16_2d", CU_TENSOR_MAP_DATA_TYPE_FLOAT16, CU_TENSOR_MAP_SWIZZLE_128B, 0, 5, 64, 32, 2, stages, f16, rows, N);
        run("f16_3d_k64", CU_TENSOR_MAP_DATA_TYPE_FLOAT16, CU_TENSOR_MAP_SWIZZLE_128B, 1, 5, 64, 64, 2, stages, f16, rows, N);
        run("f16_3d_nosw", CU_TENSOR_MAP_DATA_TYPE_FLOAT16, CU_TENSOR_MAP_SWIZZLE_NONE, 1, 5, 64, 32, 2, stages, f16, rows, N);
    }
    return 0;
}
