// Probe: 2-D tensor map over a buffer seen as rows of 4*W floats; box [48 x 6] at arbitrary column.
// nvcc -gencode arch=compute_100a,code=sm_100a -I cista-flow_b200/csrc -o build/quad_probe scripts/experiments/quad_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tma.cuh"
using namespace cf;
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int c0, int c1, float *out) {
    extern __shared__ __align__(128) uint8_t smem[];
    float *dst = reinterpret_cast<float *>(smem);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 48 * 6 * 4);
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        ptx::mbar_expect_tx(bar, 48 * 6 * 4);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(ptx::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
    }
    ptx::mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < 48 * 6; i += blockDim.x) out[i] = dst[i];
}
int main(int argc, char **argv) {
    int W = atoi(argv[1]), c0 = atoi(argv[2]), c1 = atoi(argv[3]), rows = 64;
    size_t n = (size_t)4 * W * rows;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, n * 4); cudaMalloc(&o, 48 * 6 * 4);
    cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    CUtensorMap tmap;
    cuuint64_t dims[2] = {(cuuint64_t)4 * W, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)16 * W};
    cuuint32_t box[2] = {48, 6}, estr[2] = {1, 1};
    CUresult r = tensor_map_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("W=%d c0=%d c1=%d encode=%d ", W, c0, c1, (int)r);
    probe<<<1, 128, 48 * 6 * 4 + 64>>>(tmap, c0, c1, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> g(48 * 6);
        cudaMemcpy(g.data(), o, 48 * 6 * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 48; ++j) {
            float want = (c0 + j < 4 * W && c1 + i < rows) ? (float)((size_t)(c1 + i) * 4 * W + c0 + j) : 0.f;
            bad += g[i * 48 + j] != want;
        }
        printf("mismatches=%d", bad);
    }
    printf("\n");
    return 0;
}
