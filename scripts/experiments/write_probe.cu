// Round-2 probe: what does the B200 sustain for a WRITE-ONLY stream the size of the correlation volume, and does the
// store pattern of corr_build_tc.cu's epilogue (32x32 fp32 boxes of 128x160 tiles, tensor-map bulk stores) reach it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/write_probe scripts/experiments/write_probe.cu -lcuda
// The roofline of the pyramid build is quoted against the measured COPY peak (read + write); a kernel that only writes
// may have a different ceiling.  Modes (all write B x N x N fp32, N = 4800, far more than the L2):
//   memset        cudaMemsetAsync
//   st_v4         st.global.v4 from registers, linear, persistent grid
//   st_v4_cs      the same with the .cs (streaming) hint
//   bulk_16k      cp.async.bulk shared -> global, linear 16 KB pieces, ring of R per CTA
//   box_tile      the epilogue's pattern without the GEMM: persistent CTAs walk 128x160 tiles (nb fastest, interleaved
//                 over the grid), 8 issuing warps per CTA (4 row quadrants x 2 column halves), one {32 cols x 32 rows}
//                 tensor-map box per 32-column chunk out of a 2-buffer ring per warp
//   row_tile      the same tiles as 128 plain bulk stores of one 640-byte row segment each
#include <cuda.h>
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
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void box_store(const CUtensorMap *m, const void *ssrc, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(ssrc)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void box_store_3d(const CUtensorMap *m, const void *ssrc, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(ssrc)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int K> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(K) : "memory"); }
template <int K> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(K) : "memory"); }

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
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *m, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

template <bool CS>
__global__ void __launch_bounds__(256) st_linear_kernel(float4 *dst, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        if (CS) __stcs(dst + i, v);
        else dst[i] = v;
    }
}

template <int R>
__global__ void __launch_bounds__(32) bulk_linear_kernel(uint8_t *dst, size_t bytes) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr uint32_t PIECE = 16384;
    if (threadIdx.x == 0) {
        const size_t pieces = bytes / PIECE;
        int k = 0;
        for (size_t i = blockIdx.x; i < pieces; i += gridDim.x, ++k) {
            bulk_store(dst + i * PIECE, smem + (k % R) * PIECE, PIECE);
            bulk_commit();
            bulk_wait_read<R - 1>();
        }
        bulk_wait<0>();
    }
}

// 8 issuing warps; warp w: row quadrant w & 3, column half w >> 2 (like the epilogue with ES = 2)
template <bool BOX>
__global__ void __launch_bounds__(256) tile_store_kernel(const __grid_constant__ CUtensorMap tmap, float *dst, int B, int N, int BN) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = (N + BN - 1) / BN, tiles_m = (N + 127) / 128;
    const long long total = (long long)B * tiles_m * tiles_n;
    uint8_t *mine = smem + warp * 2 * 4096;
    const int q = warp & 3, half = warp >> 2;
    const int chunks = BN / 32;                       // 32-column chunks per tile
    const int c_lo = half * ((chunks + 1) / 2), c_hi = half ? chunks : (chunks + 1) / 2;
    if (lane != 0) return;
    int k = 0;
    for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int nb = (int)(tile % tiles_n);
        const long long t = tile / tiles_n;
        const int mb = (int)(t % tiles_m), b = (int)(t / tiles_m);
        const int row0 = mb * 128 + q * 32;
        if (BOX) {
            for (int c = c_lo; c < c_hi; ++c, ++k) {
                const int col = nb * BN + c * 32;
                if (col < N && row0 < N) box_store(&tmap, mine + (k & 1) * 4096, col, b * N + row0);
                bulk_commit();
                bulk_wait_read<1>();
            }
        } else {
            // one row segment of this warp's column half per instruction
            const int col = nb * BN + c_lo * 32, width = min((c_hi - c_lo) * 32, N - col);
            for (int r = 0; r < 32; ++r) {
                if (row0 + r < N && width > 0)
                    bulk_store(dst + ((size_t)b * N + row0 + r) * N + col, mine, (uint32_t)width * 4);
                if ((r & 7) == 7) { bulk_commit(); bulk_wait_read<1>(); }
            }
        }
    }
    bulk_wait<0>();
}

// full rows: 4 issuing warps (one per row quadrant), one bulk store per 640-byte row of the tile out of a padded
// row-major staging buffer (pitch 656 B); LANES: every lane issues its own row (one warp-wide instruction), else lane 0
// issues all 32
template <bool LANES>
__global__ void __launch_bounds__(128) row_full_kernel(float *dst, int B, int N, int BN) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = (N + BN - 1) / BN, tiles_m = (N + 127) / 128;
    const long long total = (long long)B * tiles_m * tiles_n;
    const int pitch = BN * 4 + 16;
    uint8_t *mine = smem + warp * 2 * 32 * pitch;
    if (!LANES && lane != 0) return;
    int k = 0;
    for (long long tile = blockIdx.x; tile < total; tile += gridDim.x, ++k) {
        const int nb = (int)(tile % tiles_n);
        const long long t = tile / tiles_n;
        const int mb = (int)(t % tiles_m), b = (int)(t / tiles_m);
        const int row0 = mb * 128 + warp * 32, col = nb * BN;
        const uint8_t *buf = mine + (k & 1) * 32 * pitch;
        if (LANES) {
            if (row0 + lane < N) bulk_store(dst + ((size_t)b * N + row0 + lane) * N + col, buf + lane * pitch, (uint32_t)BN * 4);
        } else {
            for (int r = 0; r < 32; ++r)
                if (row0 + r < N) bulk_store(dst + ((size_t)b * N + row0 + r) * N + col, buf + r * pitch, (uint32_t)BN * 4);
        }
        bulk_commit();
        bulk_wait_read<1>();
    }
    bulk_wait<0>();
}

// The GEMM's data movement without the GEMM: warp 8 is the operand producer (fp16 maps [B*256][N], per tile two stages of
// {64 cols x 128 K-rows x 2 atoms} (fmap1 slice, skipped when A_RES: the slice of a query block stays resident) +
// {64 x 128 x 3} (fmap2 slice) into a ring of two, freed as soon as they land); warps 0-7 store the tile's level 0:
// STORE 0: nothing, 1: 32x32 boxes (8 warps, the kernel's epilogue), 2: whole 640-byte rows (4 warps, each lane its row).
// Tile order = the kernel's (nb fastest, interleaved over the grid) or, with A_RES, runs of 15 consecutive nb per CTA.
template <int STORE, bool A_RES, int LOADS>
__global__ void __launch_bounds__(320) move_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_f, const __grid_constant__ CUtensorMap tmap_c,
            const __grid_constant__ CUtensorMap tmap_c2, const __grid_constant__ CUtensorMap tmap_c5, const __grid_constant__ CUtensorMap tmap_c3,
                                                   float *dst, const void *dst_aux, int B, int N, int BN) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[2];
    constexpr int A_BYTES = 2 * 64 * 128 * 2, B_BYTES = 3 * 64 * 128 * 2, STAGE = A_BYTES + B_BYTES;
    uint8_t *stage_mem = smem, *epi = smem + 2 * STAGE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = (N + BN - 1) / BN, tiles_m = (N + 127) / 128;
    const long long total = (long long)B * tiles_m * tiles_n;
    constexpr int RUN = 15;
    if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    auto tile_at = [&](long long k) -> long long {
        if (!A_RES) return blockIdx.x + k * gridDim.x;
        return ((long long)blockIdx.x + (k / RUN) * gridDim.x) * RUN + k % RUN;
    };
    if (warp >= 8 && LOADS == 2) {
        // operand stages through the LSU: cp.async 16 B pieces into the layout the TMA would have produced (128B swizzle),
        // one commit group per stage, two stages in flight
        const unsigned short *fm = reinterpret_cast<const unsigned short *>(dst_aux);
        const int lw = warp - 8, nlw = (int)(blockDim.x >> 5) - 8;
        long long kk = 0;
        for (long long k = 0; tile_at(k) < total; ++k) {
            const long long tile = tile_at(k);
            const int nb = (int)(tile % tiles_n);
            const long long t = tile / tiles_n;
            const int mb = (int)(t % tiles_m), b = (int)(t / tiles_m);
            for (int kb = 0; kb < 2; ++kb, ++kk) {
                uint8_t *sa = stage_mem + (kk & 1) * STAGE, *sb = sa + A_BYTES;
                const size_t row0 = (size_t)b * 256 + kb * 128;
                for (int c = lw * 32 + lane; c < 2048; c += 32 * nlw) {
                    const int kr = c >> 4, cc = c & 15;
                    int col = mb * 128 + cc * 8;
                    if (col > N - 8) col = N - 8;
                    const void *src = fm + (row0 + kr) * N + col;
                    const uint32_t d = smem_u32(sa + (cc >> 3) * 16384 + kr * 128 + (((cc & 7) ^ (kr & 7)) << 4));
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
                }
                for (int c = lw * 32 + lane; c < 2560; c += 32 * nlw) {
                    const int kr = c / 20, cc = c % 20;
                    const void *src = fm + (row0 + kr) * N + nb * BN + cc * 8;
                    const uint32_t d = smem_u32(sb + (cc >> 3) * 16384 + kr * 128 + (((cc & 7) ^ (kr & 7)) << 4));
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        return;
    }
    if (warp >= 8) {
        if (warp > 8 || LOADS != 1 || lane != 0) return;
        uint32_t phase_bits = 0;
        long long issued = 0, waited = 0;
        long long my_tiles = 0;
        for (long long k = 0; tile_at(k) < total; ++k) ++my_tiles;
        const long long stages_total = my_tiles * 2;
        while (waited < stages_total) {
            while (issued < stages_total && issued - waited < 2) {
                const int s = (int)(issued & 1);
                const long long tile = tile_at(issued >> 1);
                const int kb = (int)(issued & 1);
                const int nb = (int)(tile % tiles_n);
                const long long t = tile / tiles_n;
                const int mb = (int)(t % tiles_m), b = (int)(t / tiles_m);
                const bool with_a = !A_RES || ((issued >> 1) % RUN == 0);
                mbar_expect_tx(&full[s], (uint32_t)((with_a ? A_BYTES : 0) + B_BYTES));
                if (with_a) tma_load_3d(stage_mem + s * STAGE, &tmap_a, 0, b * 256 + kb * 128, mb * 2, &full[s]);
                tma_load_3d(stage_mem + s * STAGE + A_BYTES, &tmap_f, 0, b * 256 + kb * 128, (nb * BN) / 64, &full[s]);
                ++issued;
            }
            const int s = (int)(waited & 1);
            while (!mbar_try_wait(&full[s], (phase_bits >> s) & 1u)) {}
            phase_bits ^= 1u << s;
            ++waited;
        }
        return;
    }
    if (STORE == 0) return;
    if (STORE == 1) {
        uint8_t *mine = epi + warp * 2 * 4096;
        const int q = warp & 3, half = warp >> 2;
        const int chunks = BN / 32;
        if (lane != 0) return;
        int kk = 0;
        for (long long k = 0; tile_at(k) < total; ++k) {
            const long long tile = tile_at(k);
            const int nb = (int)(tile % tiles_n);
            const long long t = tile / tiles_n;
            const int mb = (int)(t % tiles_m), b = (int)(t / tiles_m);
            const int row0 = mb * 128 + q * 32;
            for (int c = half; c < chunks; c += 2, ++kk) {
                if (row0 < N) box_store(&tmap_c, mine + (kk & 1) * 4096, nb * BN + c * 32, b * N + row0);
                bulk_commit();
                bulk_wait_read<1>();
            }
        }
        bulk_wait<0>();
    } else if (STORE == 3) {
        // 2-atom boxes {32 cols, 32 rows, 2}: half 0 stores chunks (0,1) and then chunk 4 (a single box), half 1 chunks (2,3);
        // ONE 8 KB buffer per warp (what the kernel's 64 KB of epilogue buffers hold)
        uint8_t *mine = epi + warp * 8192;
        const int q = warp & 3, half = warp >> 2;
        if (lane != 0) return;
        for (long long k = 0; tile_at(k) < total; ++k) {
            const long long tile = tile_at(k);
            const int nb = (int)(tile % tiles_n);
            const long long t = tile / tiles_n;
            const int mb = (int)(t % tiles_m), b = (int)(t / tiles_m);
            const int row0 = mb * 128 + q * 32;
            bulk_wait_read<0>();
            if (row0 < N) box_store_3d(&tmap_c2, mine, 0, b * N + row0, (nb * BN) / 32 + 2 * half);
            bulk_commit();
            if (half == 0) {
                bulk_wait_read<0>();
                if (row0 < N) box_store(&tmap_c, mine, nb * BN + 128, b * N + row0);
                bulk_commit();
            }
        }
        bulk_wait<0>();
    } else if (STORE == 5) {
        // one box per WARP and tile: the first warp of a lane quarter stores chunks 0-2 as {32, 32, 3}, the second chunks 3-4 as
        // {32, 32, 2}; one buffer per warp (12 / 8 KB), no barrier between the warps
        uint8_t *mine = epi + ((warp & 3) * 20480 + (warp >> 2) * 12288) % 49152;   // (probe: buffers may overlap, contents do not matter)
        const int q = warp & 3, half = warp >> 2;
        if (lane != 0) return;
        for (long long k = 0; tile_at(k) < total; ++k) {
            const long long tile = tile_at(k);
            const int nb = (int)(tile % tiles_n);
            const long long t = tile / tiles_n;
            const int mb = (int)(t % tiles_m), b = (int)(t / tiles_m);
            const int row0 = mb * 128 + q * 32;
            bulk_wait_read<0>();
            if (row0 < N) {
                if (half == 0) box_store_3d(&tmap_c3, mine, 0, b * N + row0, (nb * BN) / 32);
                else box_store_3d(&tmap_c2, mine, 0, b * N + row0, (nb * BN) / 32 + 3);
            }
            bulk_commit();
        }
        bulk_wait<0>();
    } else if (STORE == 6) {
        // level 0 straight from registers in the tcgen05.ld.16x256b fragment layout: thread t holds (row t/4, columns 2(t%4), +1)
        // and (row t/4 + 8, same columns) of every 8-column block -- four threads cover one 32-byte sector, a warp-wide
        // st.global.v2 eight rows x 32 B.  No shared memory, no TMA store.  Warp w: rows 32 (w & 3) .. +31, column half w >> 2.
        const int q = warp & 3, half = warp >> 2;
        const int c_lo = half ? 96 : 0, c_hi = half ? BN : 96;
        const float2 v = make_float2(1.f, 2.f);
        for (long long k = 0; tile_at(k) < total; ++k) {
            const long long tile = tile_at(k);
            const int nb = (int)(tile % tiles_n);
            const long long t = tile / tiles_n;
            const int mb = (int)(t % tiles_m), b = (int)(t / tiles_m);
            const int row0 = mb * 128 + q * 32 + (lane >> 2);
            float *base = dst + ((size_t)b * N + row0) * N + nb * BN + 2 * (lane & 3);
#pragma unroll 1
            for (int c = c_lo; c < c_hi; c += 8) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (row0 + 8 * r < N) *reinterpret_cast<float2 *>(base + (size_t)(8 * r) * N + c) = v;
            }
        }
    } else if (STORE == 4) {
        // one 5-atom box {32, 32, 5} = the whole 32 x 160 slab of a lane quarter per instruction (probe only: 20 KB staging each)
        if (warp >= 4 || lane != 0) return;
        for (long long k = 0; tile_at(k) < total; ++k) {
            const long long tile = tile_at(k);
            const int nb = (int)(tile % tiles_n);
            const long long t = tile / tiles_n;
            const int mb = (int)(t % tiles_m), b = (int)(t / tiles_m);
            const int row0 = mb * 128 + warp * 32;
            bulk_wait_read<0>();
            if (row0 < N) box_store_3d(&tmap_c5, epi, 0, b * N + row0, (nb * BN) / 32);
            bulk_commit();
        }
        bulk_wait<0>();
    } else {
        if (warp >= 4) return;
        const int pitch = BN * 4 + 16;
        uint8_t *mine = epi;      // (probe: one 32-row buffer read by all four quadrants -- the contents do not matter)
        for (long long k = 0; tile_at(k) < total; ++k) {
            const long long tile = tile_at(k);
            const int nb = (int)(tile % tiles_n);
            const long long t = tile / tiles_n;
            const int mb = (int)(t % tiles_m), b = (int)(t / tiles_m);
            const int row0 = mb * 128 + warp * 32;
            if (row0 + lane < N) bulk_store(dst + ((size_t)b * N + row0 + lane) * N + nb * BN, mine + lane * pitch, (uint32_t)BN * 4);
            bulk_commit();
            bulk_wait_read<0>();
        }
        bulk_wait<0>();
    }
}

template <typename F>
static void timed(const char *name, double bytes, F launch) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f, sum = 0.f;
    const int reps = 5;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        best = ms < best ? ms : best;
        sum += ms;
    }
    CK(cudaGetLastError());
    printf("%-22s %8.3f ms mean, %8.3f best  -> %7.1f GB/s mean, %7.1f best\n", name, sum / reps, best,
           bytes / (sum / reps * 1e-3) / 1e9, bytes / (best * 1e-3) / 1e9);
}

int main(int argc, char **argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 16, N = 4800, BN = 160;
    const size_t bytes = (size_t)B * N * N * 4;
    uint8_t *buf;
    CK(cudaMalloc(&buf, bytes));
    printf("write-only stream of %d x %d x %d fp32 = %.2f GB\n", B, N, N, bytes / 1e9);
    timed("memset", (double)bytes, [&] { CK(cudaMemsetAsync(buf, 0, bytes)); });
    for (int per_sm : {4, 8}) {
        char nm[64];
        snprintf(nm, sizeof nm, "st_v4 (%d CTAs/SM)", per_sm);
        timed(nm, (double)bytes, [&] { st_linear_kernel<false><<<148 * per_sm, 256>>>((float4 *)buf, bytes / 16); });
        snprintf(nm, sizeof nm, "st_v4_cs (%d CTAs/SM)", per_sm);
        timed(nm, (double)bytes, [&] { st_linear_kernel<true><<<148 * per_sm, 256>>>((float4 *)buf, bytes / 16); });
    }
    CK(cudaFuncSetAttribute(bulk_linear_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384));
    CK(cudaFuncSetAttribute(bulk_linear_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384));
    timed("bulk_16k ring 4, 1/SM", (double)bytes, [&] { bulk_linear_kernel<4><<<148, 32, 4 * 16384>>>(buf, bytes); });
    timed("bulk_16k ring 8, 1/SM", (double)bytes, [&] { bulk_linear_kernel<8><<<148, 32, 8 * 16384>>>(buf, bytes); });
    timed("bulk_16k ring 4, 2/SM", (double)bytes, [&] { bulk_linear_kernel<4><<<296, 32, 4 * 16384>>>(buf, bytes); });

    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)B * N};
    cuuint64_t strides[1] = {(cuuint64_t)N * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed (%d)\n", (int)r); return 1; }
    const int smem = 8 * 2 * 4096;
    CK(cudaFuncSetAttribute(tile_store_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(tile_store_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int per_sm : {1, 2}) {
        char nm[64];
        snprintf(nm, sizeof nm, "box_tile (%d CTAs/SM)", per_sm);
        timed(nm, (double)bytes, [&] { tile_store_kernel<true><<<148 * per_sm, 256, smem>>>(m, (float *)buf, B, N, BN); });
        snprintf(nm, sizeof nm, "row_tile (%d CTAs/SM)", per_sm);
        timed(nm, (double)bytes, [&] { tile_store_kernel<false><<<148 * per_sm, 256, smem>>>(m, (float *)buf, B, N, BN); });
    }
    const int smem_rows = 4 * 2 * 32 * (BN * 4 + 16);
    CK(cudaFuncSetAttribute(row_full_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_rows));
    CK(cudaFuncSetAttribute(row_full_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_rows));
    timed("row_full lane0 (1/SM)", (double)bytes, [&] { row_full_kernel<false><<<148, 128, smem_rows>>>((float *)buf, B, N, BN); });
    timed("row_full lanes (1/SM)", (double)bytes, [&] { row_full_kernel<true><<<148, 128, smem_rows>>>((float *)buf, B, N, BN); });
    // ---- loads and stores together
    {
        void *fm;
        const size_t fbytes = (size_t)B * 256 * N * 2;
        CK(cudaMalloc(&fm, fbytes));
        CK(cudaMemset(fm, 0, fbytes));
        CUtensorMap mf;
        cuuint64_t fd[3] = {64, (cuuint64_t)B * 256, (cuuint64_t)(N / 64)};
        cuuint64_t fs[2] = {(cuuint64_t)N * 2, 128};
        cuuint32_t fb3[3] = {64, 128, 3}, fb2[3] = {64, 128, 2}, e3[3] = {1, 1, 1};
        CUtensorMap mfa;
        CUresult r0 = encode_fn()(&mfa, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, fm, fd, fs, fb2, e3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r0 != CUDA_SUCCESS) { printf("fmap (A) encode failed (%d)\n", (int)r0); return 1; }
        CUresult r1 = encode_fn()(&mf, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, fm, fd, fs, fb3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r1 != CUDA_SUCCESS) { printf("fmap encode failed (%d)\n", (int)r1); return 1; }
        const int stage = (2 + 3) * 64 * 128 * 2;
        const int sm_box = 2 * stage + 8 * 2 * 4096, sm_row = 2 * stage + 32 * (BN * 4 + 16);
        printf("loads + stores together (per tile: 2 stages of 80 KB in, 80 KB out); sm_box %d sm_row %d\n", sm_box, sm_row);
        CUtensorMap mc2, mc5, mc3;
        {
            cuuint64_t cd[3] = {32, (cuuint64_t)B * N, (cuuint64_t)(N / 32)};
            cuuint64_t cs[2] = {(cuuint64_t)N * 4, 128};
            cuuint32_t b2[3] = {32, 32, 2}, b5[3] = {32, 32, 5};
            CUresult ra = encode_fn()(&mc2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, cd, cs, b2, e3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            CUresult rb = encode_fn()(&mc5, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, cd, cs, b5, e3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            cuuint32_t b3[3] = {32, 32, 3};
            CUresult rc3 = encode_fn()(&mc3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, cd, cs, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rc3 != CUDA_SUCCESS) { printf("volume 3-atom encode failed (%d)\n", (int)rc3); return 1; }
            if (ra != CUDA_SUCCESS || rb != CUDA_SUCCESS) { printf("volume 3-D encode failed (%d %d)\n", (int)ra, (int)rb); return 1; }
        }
#define RUNMOVE(NAME, ST, AR, LD, SM)                                                                                        \
        CK(cudaFuncSetAttribute(move_kernel<ST, AR, LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));                   \
        timed(NAME, (double)bytes, [&] { move_kernel<ST, AR, LD><<<148, (LD == 2 ? 320 : 288), SM>>>(mfa, mf, m, mc2, mc5, mc3, (float *)buf, fm, B, N, BN); });
        RUNMOVE("loads only", 0, false, 1, sm_box)
        RUNMOVE("loads only, A resident", 0, true, 1, sm_box)
        RUNMOVE("box stores only", 1, false, 0, sm_box)
        RUNMOVE("loads + box stores", 1, false, 1, sm_box)
        RUNMOVE("A-res loads + box", 1, true, 1, sm_box)
        RUNMOVE("LSU loads only (2 warps)", 0, false, 2, sm_box)
        RUNMOVE("LSU loads + box stores", 1, false, 2, sm_box)
        RUNMOVE("2-atom boxes only", 3, false, 0, sm_box)
        RUNMOVE("loads + 2-atom boxes", 3, false, 1, sm_box)
        RUNMOVE("fragment st.v2 only", 6, false, 0, sm_box)
        RUNMOVE("loads + fragment st.v2", 6, false, 1, sm_box)
        RUNMOVE("per-warp 3|2-atom only", 5, false, 0, sm_box)
        RUNMOVE("loads + per-warp 3|2", 5, false, 1, sm_box)
        RUNMOVE("5-atom boxes only", 4, false, 0, sm_box)
        RUNMOVE("loads + 5-atom boxes", 4, false, 1, sm_box)
        if (sm_row <= 227 * 1024) {
            RUNMOVE("row stores only", 2, false, 0, sm_row)
            RUNMOVE("loads + row stores", 2, false, 1, sm_row)
            RUNMOVE("A-res loads + row", 2, true, 1, sm_row)
        } else printf("row staging does not fit beside two 80 KB stages (%d B)\n", sm_row);
        CK(cudaFree(fm));
    }
    CK(cudaFree(buf));
    return 0;
}
