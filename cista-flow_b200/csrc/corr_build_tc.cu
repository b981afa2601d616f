// Tensor-core all-pairs correlation for sm_100a:
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring ->
//   tcgen05.mma kind::tf32 (accumulators in TMEM) -> tcgen05.ld epilogue with
//   fused 1/sqrt(D) scale and fused level-1 (2x2) average pooling.
//
// Replaces the torch.matmul + first avg_pool2d of CorrBlock.__init__
// (ERAFT/corr.py:19-27,52-60; DCEIFlow/core/corr/raft_corr.py:22-30,56-65).
//
// GEMM view, per batch item b:   C[i, j] = sum_k A[k, i] * Bm[k, j]
//   A = fmap1[b] (D x N), Bm = fmap2[b] (D x N), N = h*w contiguous: BOTH
//   operands are "MN-major" (the reference's fmap1.transpose(1,2) is free).
//   The tiles are loaded as TMA boxes of 32 (MN, 128 bytes) x 32 (K rows) with
//   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, which lands them in the only shared
//   memory layout tcgen05 accepts for MN-major 32-bit operands: UMMA layout type
//   1, "128B swizzle with 32-byte atoms" (Swizzle<2,5,2>: the 32-byte chunk index
//   of a 128-byte row is XORed with row % 4).  Canonical form, in 16-byte units:
//   ((8,n),(4,k)) : ((1,LBO),(8,SBO)) -- LBO = 4096 B between 32-column atoms
//   (one TMA box each), SBO = 512 B between groups of 4 K rows.  One tcgen05.mma
//   kind::tf32 has K = 8 = two 4-row groups, so the descriptor start address
//   advances by 1024 B per MMA.  (The plain SWIZZLE_128B layout is silently
//   computed as zeros for MN-major tf32; scripts/tc_probe.cu is the single-tile
//   probe that established this on a B200.)
//
// Tile: 128 queries (M, TMEM lanes) x BN targets (TMEM columns).  With pooling
// fused, BN = R*w covers R (even) whole rows of the target map so that each 2x2
// pooling window lives inside one accumulator row; the epilogue thread that owns
// query i reads row y and row y+1 of its TMEM lane, stores both level-0 rows and
// the pooled level-1 row.  Level 0 (N^2 floats) is therefore written once and
// never re-read: the volume is HBM-write bound (96 FLOP/B).
//
// Persistent, warp-specialised CTA (1 per SM, 192 threads):
//   warps 0-3  epilogue (TMEM lane quarter = warp id): TMEM -> registers -> swizzled smem -> TMA store
//   warp  4    TMA producer (one elected lane)
//   warp  5    TMEM allocator + MMA issuer (one elected lane)
// Two TMEM accumulator stages (2 x 256 columns) let the MMA of tile t+1 overlap
// the epilogue of tile t; a 4-stage shared-memory ring feeds the MMA.
//
// Operand precision: tcgen05 kind::tf32 reads fp32 words and ignores the low 13
// mantissa bits (truncation), which biases every product low (measured: signed
// relative bias -7e-4, max|err| 7.7e-4*max|ref| at D=256).  Encoding the tensor
// maps with CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 makes the TMA unit round the
// operands to TF32 (nearest) on their way into shared memory -- measured
// identical to a cvt.rna.tf32.f32 pre-pass: bias -5e-7, max|err| 2.8e-4*max|ref|
// -- so no extra pass over the feature maps and no workspace are needed.
//
// State at the end of round 2 (the notes below and at launch_tc() are the history that led here): on wide maps the default
// is kind::f16 on fp16 operand copies, two 128-K-row stages per tile (fmap1 slice: one 128B-swizzled 3-D box; fmap2 slice:
// one 64B-swizzled 3-D box of exactly BN columns), eight epilogue warps, levels 1 and 2 pooled in the epilogue.  At 64 x
// 60x80 the GEMM kernel runs 3.4 us per 128x160 tile = 1.65 ms (ncu) and writes at 4.7-5.0 TB/s.  Its period is the SM's
// TMA unit: four operand boxes (~0.32 us each) + twenty 4 KB store boxes (~0.12 us each) per tile queue in ONE unit
// (scripts/experiments/write_probe.cu); the epilogue warps are busy ~3.0 us of it, the MMAs ~1.5 us.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace cf {

namespace tc {

constexpr int BM = 128;          // queries per tile (UMMA M)
constexpr int BK = 32;           // K rows per pipeline stage (4 MMAs of K=8)
// fp16 operands: 64 K rows per stage.  One SM's TMA unit retires a 3-D box instruction of these shapes about every
// 320 ns WHATEVER its size between 20 and 40 KB (profiles/r02/tma_probe.txt: {64 cols, 32 rows, 5 atoms} and
// {64, 64, 5} both 320 ns), and a stage is two instructions (fmap1 box + fmap2 box): with 32-row stages the fp16
// kernel spent 8 x 2 x 320 ns = 5.1 us per tile on operand loads alone -- exactly what it measured, and why halving
// the operand bytes had bought nothing in round 1.  64-row stages: 3.0 us of loads per tile (ablation: loads only), 305 us at
// 8 x 60x80; 128-row stages (two per tile, 80 KB each, ring of two): loads + MMAs 2.45 us per tile, 290 us.
constexpr int BK_F16 = 128;
constexpr int BK_TF32 = 32;      // TF32 stages.  64-row stages (ring of two) measured the same within 2 % at every shape (64 x 60x80 2595
                                 // against 2580 us): the TF32 kernel is bound by its MMAs (3.0 us per 128x160 tile), not by TMA issue
constexpr int MAX_STAGES = 8;     // ring depth = as many stages of (fmap1 tile + this shape's fmap2 tile) as fit, at most 6
constexpr int MAX_BN = 256;      // UMMA N limit
constexpr int BOX_BYTES = 32 * 32 * 4;                 // one TMA box: 32 cols x 32 rows fp32
constexpr int A_BYTES = (BM / 32) * BOX_BYTES;         // 16 KB
constexpr int EPI_BUF_BYTES = 32 * 32 * 4;               // one 32x32 fp32 output box, 128B swizzle
// Epilogue warps per TMEM lane quarter.  2 (eight epilogue warps splitting the tile's work items) was measured
// SLOWER on the B200 (8 x 60x80: 362 against 334 us; with loads and stores ablated 258 against 238 us): the
// epilogue is not the critical path, the tf32 MMAs are (operand fetch from shared memory).
// (the kernel's template parameter ES; the TF32 kernel uses 1.  With fp16 operands the stages are half the size, the
//  64 KB of epilogue buffers fit beside a deep ring, and the epilogue IS the critical path: one tile's epilogue takes
//  3.7 us on four warps while the MMAs of a 128x160 tile need 1.5 us -- see the kernel's header.)
constexpr int SMEM_LIMIT = 227 * 1024;
__host__ __device__ constexpr int epi_bytes(int es) { return 4 * es * 2 /*double buffer*/ * EPI_BUF_BYTES; }   // 32 KB per ES
__host__ __device__ constexpr int smem_fixed(int es) { return epi_bytes(es) + 1024 /*align slack*/ + 256 /*barriers*/; }
__host__ __device__ constexpr int threads(int es) { return 32 * (4 * es + 2); }
constexpr uint32_t SPIN_LIMIT = 1u << 26;  // ~seconds; a stuck pipeline traps instead of hanging the GPU

struct Params {
    int B, D, N, h, w;
    int BN;        // valid target columns per tile
    int BN_mma;    // UMMA N (BN rounded up to 16)
    int n_boxes_b; // TMA boxes per stage for the B operand
    int R;         // target rows per tile when pooling is fused (0 = not fused)
    int tiles_m, tiles_n, total_tiles;
    int tiles_nv;  // target tiles per query block in the tile index (= tiles_n, rounded up to 4 with quad3: indices with
                   // nb >= tiles_n are skipped by every role)
    int quad3;     // 1 (with pair2): a CTA takes the tiles of FOUR consecutive target-row pairs back to back and the epilogue pools
                   // level 3 from two level-2 rows (the first one parked in free TMEM columns, like pair2's level-1 row)
    int tiles_mp;  // pair kernel: pairs of query tiles per batch item (= ceil(tiles_m / 2)); total_tiles counts pairs
    int stages;      // shared-memory ring depth
    int stage_bytes; // A_BYTES + n_boxes_b * BOX_BYTES
    int qbox_h0;     // qbox: level-0 chunks of the first warp of a lane quarter
    int frag_stores; // 1: level 0 leaves straight from registers in the tcgen05.ld.16x256b fragment layout (32-byte runs, no shared
                     // memory, no TMA store) -- see launch_tc()
    int qbox;        // 1 (two epilogue warps per lane quarter): level 0 leaves as ONE {32 cols, 32 rows, BN/32 atoms} box per lane
                     // quarter and tile (tmap_c is then that 3-D map over {32, B*N rows, N/32 atoms}) -- see launch_tc()
    int epi_bytes;   // bytes of epilogue buffers between the operand ring and the barriers
    int l1_scratch;  // byte offset (from the epilogue buffers) of a level-1 scratch area of 2 KB per epilogue warp that is NOT a
                     // level-0 box buffer (0: none, the strips borrow a box buffer and wait for its store)
    int b_sw64;      // fp16 operands, BN_mma an odd multiple of 32: the fmap2 slice as {32 cols, K rows, BN_mma / 32 atoms} with the
                     // 64-byte swizzle -- exactly BN_mma columns per stage instead of the next multiple of 64 (160: 40 KB, not 48)
    int ablate;    // experiments (CF_TC_FLAGS bits 8-10): 1 = no level-0/1 stores, 2 = no MMAs, 4 = no operand loads, 16 = no epilogue at all, 32 = no pooling part, 64 = no level-0 part
    int b_half;    // pair kernel: fmap2 boxes per stage and CTA (half of the tile's columns each)
    int h1, w1;    // level-1 map size
    float scale;
    int stream_l0; // 1: level 0 is larger than L2 can hold -> evict-first stores
    int b3d;       // as atoms3d, for the B operand: also needs tile origins on atom boundaries (BN % 32 == 0)
    int atoms3d;   // 1: N % 32 == 0 -> the tensor maps are 3-D {32 cols, rows, 32-column atoms}: ONE TMA instruction
                   //    per operand and stage.  One thread issues a cp.async.bulk.tensor about every 100 cycles, and
                   //    9-12 per-atom boxes per stage made that issue rate the bound of the main loop (timeline)
    float *l0;
    float *l1;
    int l1_staged;           // 1: level-1 rows are transposed through shared memory into 64-byte runs (w1 % 4 == 0)
    int lsu_stores;          // 1: level-0 rows leave through the LSU (transposed in shared memory), 0: TMA bulk stores
    const float *inv_scale;  // F16 operands: [2*B] powers of two that undo the per-item operand scaling (else nullptr)
    // deep fusion (tiles of 8k whole target rows, i.e. feature maps up to 32 wide -- the 180x240 / DAVIS240 case):
    // levels 2 and 3 are pooled from the level-1 rows while they are still in registers
    int pair2;     // 1: a CTA takes tiles in PAIRS of consecutive target-row pairs (nb, nb+1) and the epilogue pools level 2 from
                   //    the two level-1 rows (the first one parked in free TMEM columns) -- maps too wide for `deep`, R == 2
    int deep;      // 0: none, 1: level 2, 2: levels 2 and 3
    int h2, w2, h3, w3;
    float *l2;
    float *l3;
};

// ---- PTX wrappers ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// (L2 policy hints -- evict_last on these loads, evict_first on the volume stores -- were measured on the B200:
//  no effect at any shape, so the plain forms are used)
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *m, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *m, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// cta_group::2 variants: the box lands in THIS CTA's shared memory, its bytes complete on the mbarrier at the
// shared::cluster address `bar_addr` -- the pair leader's `full` barrier (mapa_rank0) for both CTAs of the pair
__device__ __forceinline__ void tma_load_2d_2sm(void *dst, const CUtensorMap *m, int c0, int c1, uint32_t bar_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void *dst, const CUtensorMap *m, int c0, int c1, int c2, uint32_t bar_addr) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// shared::cluster address of the same shared-memory object in CTA 0 of the cluster
__device__ __forceinline__ uint32_t mapa_rank0(const void *p) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(0u));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// smem (32 rows x 128 B, 128B swizzle) -> global box {32 cols, 32 rows, 1} of the [B][N][N] volume
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *m, const void *src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// the two epilogue warps of TMEM lane quarter q (64 threads) meet
__device__ __forceinline__ void quarter_barrier(int q) { asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory"); }
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N committed store groups may still be READING their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// pair kernel: arrive on the mbarrier at this offset in every CTA of `mask` once the pair's MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t *slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// M = 256 across the CTA pair: each CTA contributes its 128 fmap1 rows and HALF of the fmap2 tile from its own
// shared memory (same offsets in both CTAs) and receives its 128 x N accumulator rows in its own TMEM
__device__ __forceinline__ void umma_tf32_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 operands, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 consecutive columns: thread t <- lane (base+t), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 16 consecutive columns, registers -> TMEM / TMEM -> registers (the level-1 row parked between two tiles)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 16 lanes x 32 columns in the 16x256b fragment layout: thread t <- for every 8-column block j: (lane t/4, columns 8j + 2(t%4), +1) in
// v[4j], v[4j+1] and (lane t/4 + 8, same columns) in v[4j+2], v[4j+3] -- four threads hold one 32-byte run of a row
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, MN-major, 128B swizzle with 32B atoms (see file header)
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo = 4096u) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);        // start address            bits [0,14)
    d |= (uint64_t)(lbo >> 4) << 16;                 // leading byte offset (MN) bits [16,30): bytes per TMA box
    d |= (uint64_t)(512u >> 4) << 32;                // stride byte offset (K)   bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)1 << 61;                          // layout type 1 = SWIZZLE_128B_BASE32B
    return d;
}
// UMMA shared-memory descriptor, MN-major 16-bit operands, plain 128B swizzle: atoms of 64 columns (128 B) x 8 K rows
// (1024 B); LBO = bytes between 64-column atoms (one TMA box), SBO = bytes between groups of 8 K rows
__device__ __forceinline__ uint64_t make_desc_mn_f16(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(4096u >> 4) << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
// D=f32, A=B=f16, both MN-major, M=128, N=n
__host__ __device__ inline uint32_t make_idesc_f16(int n) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// instruction descriptor: D=f32, A=B=tf32, both MN-major, M=m (128, or 256 across a CTA pair), N=n
__host__ __device__ inline uint32_t make_idesc_tf32(int n, int m = BM) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

// Debug timeline of CTA 0's first tile (clock64 stamps), read back with cf_debug_tc_timeline().
__device__ unsigned long long g_tc_timeline[32];
__device__ __forceinline__ void stamp(int slot) {
    if (blockIdx.x == 0) g_tc_timeline[slot] = (unsigned long long)clock64();
}

// experiment builds (-DCF_TRACE, scripts/corr_trace.py): per-CTA, per-tile globaltimer stamps, 16 slots per tile:
//  0 first load issued, 1 last load issued | 2 accumulator free (MMA), 3..10 stage kb landed, 11 last commit |
//  12 accumulator ready (epilogue), 13 level-0 chunks issued, 14 epilogue done
#ifdef CF_TRACE
#define TC_TRACE(tile_no, slot) do { if ((tile_no) < 15) CF_TRACE_AT(16 + 16 * (tile_no) + (slot)); } while (0)
#else
#define TC_TRACE(tile_no, slot) ((void)0)
#endif

__device__ __forceinline__ void decode_tile(const Params &p, int tile, int &b, int &mb, int &nb) {
    nb = tile % p.tiles_nv;
    const int t = tile / p.tiles_nv;
    mb = t % p.tiles_m;
    b = t / p.tiles_m;
}
// pair kernel: `tile` indexes a pair of vertically adjacent query tiles; CTA `rank` owns query tile 2*pair + rank,
// which may lie entirely below the volume when tiles_m is odd (its loads are zero-filled, its stores clipped)
__device__ __forceinline__ void decode_pair_tile(const Params &p, int tile, int rank, int &b, int &mb, int &nb) {
    nb = tile % p.tiles_n;
    const int t = tile / p.tiles_n;
    mb = 2 * (t % p.tiles_mp) + rank;
    b = t / p.tiles_mp;
}

template <bool STREAM>
__device__ __forceinline__ void store_row_chunk_impl(float *dst, const uint32_t (&v)[32], float scale, int nvalid, bool vec4) {
    // nvalid in [0,32], multiple of 4 when vec4
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (vec4) {
            if (4 * q < nvalid) {
                const float4 o = make_float4(__uint_as_float(v[4 * q]) * scale, __uint_as_float(v[4 * q + 1]) * scale,
                                             __uint_as_float(v[4 * q + 2]) * scale, __uint_as_float(v[4 * q + 3]) * scale);
                if (STREAM) __stcs(reinterpret_cast<float4 *>(dst) + q, o);
                else *(reinterpret_cast<float4 *>(dst) + q) = o;
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (4 * q + r < nvalid) {
                    const float o = __uint_as_float(v[4 * q + r]) * scale;
                    if (STREAM) __stcs(dst + 4 * q + r, o);
                    else dst[4 * q + r] = o;
                }
        }
    }
}
__device__ __forceinline__ void store_row_chunk(float *dst, const uint32_t (&v)[32], float scale, int nvalid, bool vec4, bool stream) {
    if (stream) store_row_chunk_impl<true>(dst, v, scale, nvalid, vec4);
    else store_row_chunk_impl<false>(dst, v, scale, nvalid, vec4);
}

// CL = 1: one CTA per tile.  CL = 2 (tcgen05 cta_group::2): a cluster of two CTAs owns two vertically adjacent
// query tiles of the same target tile = one 256 x BN UMMA.  Each CTA loads its own 128 fmap1 rows and HALF of the
// fmap2 columns; the leader (rank 0) issues the MMAs for the pair, each CTA's accumulator rows land in its own TMEM
// and are drained by its own epilogue.  Per SM and K step this moves and reads (128 + BN/2) x 32 B instead of
// (128 + BN) x 32 B: the single-CTA tf32 MMA is bound by operand fetch from shared memory (measured: ~50 B/clk,
// i.e. 181 cycles for 128x160x8 and 243 for 128x256x8 with loads and stores ablated; K-major descriptors on the
// same bytes change nothing).  Barriers: both producers' bytes complete on the LEADER's `full`; the leader's
// commits arrive on `empty` / `tfull` of both CTAs; both epilogues arrive on the leader's `tempty`.
// (A first pair variant only multicast the fmap2 tile to two independent cta_group::1 MMAs: L2 traffic -33 %,
//  time unchanged -- L2 bandwidth was not the bound.)
// F16: the operands are fp16 copies of the feature maps (same 11-bit significand as TF32, scaled per batch item by a
// power of two so that any finite input fits; made by fmap_to_half_kernel).  kind::f16 covers K = 16 per MMA where
// kind::tf32 covers 8 at the same ~1 accumulator column per clock, and every stage moves half the bytes: the GEMM
// stops being MMA / load-latency bound and runs at the HBM write rate of the volume.
template <int CL, bool F16, int ES, int BKT>
__global__ void __launch_bounds__(threads(ES), 1)
corr_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    static_assert(!(F16 && CL == 2), "the CTA-pair variant is tf32 only");
    constexpr int EPI_SPLIT = ES, EPI_WARPS = 4 * ES, EPI_BYTES = epi_bytes(ES);
#ifdef CF_SWAP_SERVICE_WARPS   // experiment: which warp scheduler (warp % 4) hosts the producer / the MMA issuer
    constexpr int PRODUCER_WARP = EPI_WARPS + 1, MMA_WARP = EPI_WARPS;
#else
    constexpr int PRODUCER_WARP = EPI_WARPS, MMA_WARP = EPI_WARPS + 1;
#endif
    constexpr int BC = F16 ? 64 : 32;                   // operand columns per TMA box (128 bytes)
    constexpr int BKK = F16 ? BK_F16 : BKT;             // K rows per stage (TF32: 64, or 32 for D % 64 != 0 and the CTA-pair kernel)
    constexpr int BOXB = 128 * BKK;                     // bytes of one TMA box: 128-byte rows x BKK
    constexpr int ABYTES = (BM / BC) * BOXB;            // fmap1 part of a stage
    constexpr int MMAS = F16 ? BKK / 16 : BKK / 8;      // MMAs per stage
    constexpr int KSTEP = F16 ? 2048 : 1024;            // descriptor start-address advance per MMA
    const int STAGES = p.stages, STAGE_BYTES = p.stage_bytes;
    uint8_t *epi = smem + STAGES * STAGE_BYTES;  // [EPI_WARPS][2][EPI_BUF_BYTES]
    uint64_t *bars = reinterpret_cast<uint64_t *>(epi + p.epi_bytes);
    uint64_t *full = bars, *empty = bars + MAX_STAGES;
    uint64_t *tfull = bars + 2 * MAX_STAGES, *tempty = bars + 2 * MAX_STAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = p.D / BKK;
    const int rank = CL == 2 ? (int)cluster_ctarank() : 0;
    const int first_tile = (int)blockIdx.x / CL, tile_step = (int)gridDim.x / CL;
    // the s-th tile of this CTA: tile indices have the target-row pair nb fastest, so with pair2 a CTA takes the two tiles
    // 2u, 2u+1 of unit u back to back (same batch item, same query rows, target rows 4u' .. 4u'+3)
    const int GS = p.quad3 ? 4 : (p.pair2 ? 2 : 1), gs_shift = p.quad3 ? 2 : (p.pair2 ? 1 : 0);
    auto tile_of = [&](int sq) { return ((first_tile + (sq >> gs_shift) * tile_step) << gs_shift) + (sq & (GS - 1)); };
    // (b, mb, nb) of the next tile of this CTA without integer divisions (three of them per tile, twice with the look-ahead
    // of the fp16 epilogue, were ~0.5 us of every tile's serial chain): the step from the last tile of a group to the first
    // of the next is a constant, decomposed once
    const int jump = GS * tile_step - (GS - 1);
    const int jump_n = jump % p.tiles_nv, jump_m = (jump / p.tiles_nv) % p.tiles_m, jump_b = (jump / p.tiles_nv) / p.tiles_m;
    auto advance = [&](int sq, int &tile, int &b, int &mb, int &nb) {   // tile of sequence number sq -> sq + 1   (CL == 1)
        if ((sq & (GS - 1)) != GS - 1) {   // next tile of the group: the next target-row pair (tiles_nv is a multiple of GS)
            ++tile;
            ++nb;
            return;
        }
        tile += jump;
        nb += jump_n;
        int c = nb >= p.tiles_nv ? 1 : 0;
        nb -= c ? p.tiles_nv : 0;
        mb += jump_m + c;
        c = mb >= p.tiles_m ? 1 : 0;
        mb -= c ? p.tiles_m : 0;
        b += jump_b + c;
    };
    if (threadIdx.x == 0) stamp(0);

    if (warp == PRODUCER_WARP && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        tma_prefetch_desc(&tmap_c);
    }
    if (warp == MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 32 * EPI_WARPS * CL); }
            fence_barrier_init();
        }
        __syncwarp();
        if (CL == 2) tmem_alloc_2sm(tmem_slot, 512);
        else tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();  // the peer's barriers are initialised (and its TMEM allocated) before anything targets them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) stamp(1);

    if (warp == PRODUCER_WARP) {
        // ------------------------------------------------------------ TMA producer
        // (whole warp convergent, an elected lane issues -- see the MMA issuer below: under `if (lane == 0)` every
        //  cp.async.bulk.tensor cost an ELECT / R2UR / BRA.U.ANY sequence, the "one TMA issue per ~100 cycles" measured
        //  earlier)
        {
            int stage = 0;
            uint32_t phase = 0;
            // pair kernel: the leader's barrier counts the bytes of BOTH CTAs' boxes
            const uint32_t tx_bytes = (F16 && p.b_sw64) ? (uint32_t)p.stage_bytes
                                                       : (uint32_t)(CL * (BM / BC + (CL == 2 ? p.b_half : p.n_boxes_b))) * BOXB;
            const int jr = CL == 2 ? rank * (p.BN_mma / 2) : 0;  // first fmap2 column of this CTA inside the tile
            int tile_no = 0;
            int tile = tile_of(0), b = 0, mb = 0, nb = 0;
            if (CL == 1 && tile < p.total_tiles) decode_tile(p, tile, b, mb, nb);
            for (; tile < p.total_tiles; ++tile_no) {
                if (CL == 2) decode_pair_tile(p, tile, rank, b, mb, nb);
                if (CL == 1 && nb >= p.tiles_n) {   // (quad3: padding index, no such tile)
                    advance(tile_no, tile, b, mb, nb);
                    continue;
                }
                const int i0 = mb * BM, j0 = nb * p.BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (lane == 0) {
                        if (kb == 0) TC_TRACE(tile_no, 0);
                        if (kb == kblocks - 1) TC_TRACE(tile_no, 1);
                    }
                    uint8_t *sa = smem + stage * STAGE_BYTES, *sb = sa + ABYTES;
                    const int krow = b * p.D + kb * BKK;
                    const bool issuer = elect_one();
                    if (!issuer) {
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    if (p.ablate & 4) {
                        if (rank == 0) mbar_arrive(&full[stage]);
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    if (CL == 2) {
                        if (rank == 0) mbar_expect_tx(&full[stage], tx_bytes);
                        const uint32_t lead_full = mapa_rank0(&full[stage]);
                        if (p.atoms3d) {
                            tma_load_3d_2sm(sa, &tmap_a, 0, krow, i0 >> 5, lead_full);
                        } else {
#pragma unroll
                            for (int a = 0; a < BM / 32; ++a) tma_load_2d_2sm(sa + a * BOX_BYTES, &tmap_a, i0 + 32 * a, krow, lead_full);
                        }
                        if (p.b3d) {
                            tma_load_3d_2sm(sb, &tmap_b, 0, krow, (j0 + jr) >> 5, lead_full);
                        } else {
                            for (int a = 0; a < p.b_half; ++a) tma_load_2d_2sm(sb + a * BOX_BYTES, &tmap_b, j0 + jr + 32 * a, krow, lead_full);
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_expect_tx(&full[stage], tx_bytes);
                    if (p.atoms3d) {
                        tma_load_3d(sa, &tmap_a, 0, krow, i0 / BC, &full[stage]);
                    } else {
#pragma unroll
                        for (int a = 0; a < BM / BC; ++a) tma_load_2d(sa + a * BOXB, &tmap_a, i0 + BC * a, krow, &full[stage]);
                    }
                    if (F16 && p.b_sw64) {
                        tma_load_3d(sb, &tmap_b, 0, krow, j0 / 32, &full[stage]);
                    } else if (p.b3d) {
                        tma_load_3d(sb, &tmap_b, 0, krow, j0 / BC, &full[stage]);
                    } else {
                        for (int a = 0; a < p.n_boxes_b; ++a) tma_load_2d(sb + a * BOXB, &tmap_b, j0 + BC * a, krow, &full[stage]);
                    }
                    if (tile_no == 0 && kb == 0) stamp(2);
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (CL == 2) tile = tile_of(tile_no + 1);
                else advance(tile_no, tile, b, mb, nb);
            }
        }
    } else if (warp == MMA_WARP) {
        // -------------------------------------------------------------- MMA issuer
        // The WHOLE warp walks the loop (convergent, every value warp-uniform) and one elected lane issues: written as
        // `if (lane == 0) { ... }` the compiler wrapped every tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY
        // sequence and rebuilt both 64-bit descriptors in front of it -- ~30 dependent uniform-datapath instructions
        // per MMA, i.e. an issue interval of ~180 cycles for an MMA that occupies the tensor pipe for ~90.
        if (rank == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            const uint32_t idesc = F16 ? make_idesc_f16(p.BN_mma) : make_idesc_tf32(p.BN_mma, CL * BM);
            // descriptor high words are constants of the layout; the low word = start address >> 4 | LBO << 16
            const uint32_t desc_hi = F16 ? (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29))
                                         : (uint32_t)((512u >> 4) | (1u << 14) | (1u << 29));
            const uint32_t lbo_bits = ((uint32_t)BOXB >> 4) << 16;   // bytes between column atoms = one TMA box
            // fmap2 slice in the 64-byte swizzle (b_sw64): atoms of 32 columns x 64 B rows, 8-row groups 512 B apart, one
            // atom = BKK x 64 B, and a K step of 16 rows is 1024 B
            const bool sw64 = F16 && p.b_sw64;
            const uint32_t desc_hi_b = sw64 ? (uint32_t)((512u >> 4) | (1u << 14) | (4u << 29)) : desc_hi;
            const uint32_t lbo_bits_b = sw64 ? (((uint32_t)(BKK * 64) >> 4) << 16) : lbo_bits;
            const uint32_t kstep_b = sw64 ? (1024u >> 4) : (uint32_t)(KSTEP >> 4);
            const uint32_t smem_base = smem_u32(smem);
            int tile_no = 0;
            int wb_ = 0, wmb_ = 0, wnb_ = 0, wtile_ = tile_of(0);          // (b, mb, nb) carried only to skip quad3's padding indices
            if (CL == 1 && wtile_ < p.total_tiles) decode_tile(p, wtile_, wb_, wmb_, wnb_);
            for (int tile = tile_of(0); tile < p.total_tiles; tile = tile_of(++tile_no)) {
                if (CL == 1) {
                    const bool skip = wnb_ >= p.tiles_n;
                    advance(tile_no, wtile_, wb_, wmb_, wnb_);
                    if (skip) continue;
                }
                mbar_wait(&tempty[acc], acc_phase ^ 1);  // epilogue drained this accumulator
                tc_fence_after();
                if (lane == 0) TC_TRACE(tile_no, 2);
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * MAX_BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        if (tile_no == 0 && kb < 16) stamp(3 + kb);
                        if (kb < 8) TC_TRACE(tile_no, 3 + kb);
                    }
                    const uint32_t sa = smem_base + (uint32_t)(stage * STAGE_BYTES);
                    const uint32_t a_lo = (((sa & 0x3FFFFu) >> 4) | lbo_bits), b_lo = ((((sa + ABYTES) & 0x3FFFFu) >> 4) | lbo_bits_b);
                    if (elect_one()) {
                        if (!(p.ablate & 2)) {
#pragma unroll
                            for (int kk = 0; kk < MMAS; ++kk) {
                                const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + (uint32_t)(kk * (KSTEP >> 4)));
                                const uint64_t db = ((uint64_t)desc_hi_b << 32) | (uint64_t)(b_lo + (uint32_t)kk * kstep_b);
                                if (F16) umma_f16(d_tmem, da, db, idesc, (uint32_t)((kb | kk) != 0));
                                else if (CL == 2) umma_tf32_2sm(d_tmem, da, db, idesc, (uint32_t)((kb | kk) != 0));
                                else umma_tf32(d_tmem, da, db, idesc, (uint32_t)((kb | kk) != 0));
                            }
                        }
                        // frees the smem slot once these MMAs have read it (pair kernel: in both CTAs)
                        if (CL == 2) umma_commit_2sm(&empty[stage], (uint16_t)3);
                        else umma_commit(&empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) {
                    if (CL == 2) umma_commit_2sm(&tfull[acc], (uint16_t)3);  // accumulator complete -> both epilogues
                    else umma_commit(&tfull[acc]);
                }
                __syncwarp();
                if (lane == 0) {
                    if (tile_no == 0) stamp(19);
                    TC_TRACE(tile_no, 11);
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue
        // warp <-> TMEM lane quarter warp % 4 (a warp may only read those lanes); with EPI_SPLIT > 1 the warps of a
        // quarter take the tile's work items (32-column level-0 chunks, 32-column pooling strips) in turn
        int acc = 0;
        uint32_t acc_phase = 0, chunk_count = 0;
        const int quarter = warp & 3, half = warp >> 2;
        const int row = 32 * quarter + lane;
        const bool vec4_l0 = (p.N % 4) == 0 && (p.w % 4) == 0;
        const bool vec4_l1 = (p.w1 % 4) == 0;
        uint8_t *my_epi = epi + warp * 2 * EPI_BUF_BYTES;
        int tile_no = 0;
        // F16: 1/sqrt(D) times the powers of two that undo the per-item scaling of the operands.  The two factors of the NEXT
        // tile's item are fetched while this tile is drained: the volume streams through the L2, so these loads are DRAM
        // misses, and fetched at the top of their own tile they put ~0.9 us into every tile of the slowest epilogue warp --
        // which sets the tile period (corr_trace, round 2)
        float inv_a = 1.f, inv_b = 1.f;
        int tile = tile_of(0), b = 0, mb = 0, nb = 0;
        if (CL == 1 && tile < p.total_tiles) decode_tile(p, tile, b, mb, nb);
        int tile_nx = tile, b_nx = b, mb_nx = mb, nb_nx = nb;      // the tile after this one (CL == 1)
        if (CL == 1) advance(0, tile_nx, b_nx, mb_nx, nb_nx);
        if (F16 && tile < p.total_tiles) {
            inv_a = __ldg(p.inv_scale + b);
            inv_b = __ldg(p.inv_scale + p.B + b);
        }
        auto next_tile = [&]() {
            ++tile_no;
            if (CL == 2) { tile = tile_of(tile_no); return; }
            tile = tile_nx; b = b_nx; mb = mb_nx; nb = nb_nx;
            advance(tile_no, tile_nx, b_nx, mb_nx, nb_nx);
        };
        for (; tile < p.total_tiles; next_tile()) {
            if (CL == 2) decode_pair_tile(p, tile, rank, b, mb, nb);
            const int i = mb * BM + row;
            const bool row_ok = i < p.N;
            const float scale = F16 ? p.scale * inv_a * inv_b : p.scale;
            if (F16 && tile_nx < p.total_tiles) {
                inv_a = __ldg(p.inv_scale + b_nx);
                inv_b = __ldg(p.inv_scale + p.B + b_nx);
            }
            if (CL == 1 && nb >= p.tiles_n) continue;   // (quad3: padding index -- after the look-ahead, which every index owes its successor)
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            if (tile_no == 0 && threadIdx.x == 0) stamp(20);
            if (threadIdx.x == 0) TC_TRACE(tile_no, 12);
            if (p.ablate & 16) {  // experiment: the epilogue reads nothing (MMA issue rate without TMEM read traffic)
                tc_fence_before();
                if (CL == 2) mbar_arrive_cluster(mapa_rank0(&tempty[acc]));
                else mbar_arrive(&tempty[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
                continue;
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)acc * MAX_BN;
            const int j0 = nb * p.BN;
            const int bn_valid = min(p.BN, p.N - j0);  // accumulator column c <-> target index j0 + c
            // ---- level 0: TMEM -> registers (x scale) -> swizzled smem box -> TMA bulk store.
            // A warp-wide st.global of this fragment would touch 32 different rows per instruction
            // (measured: 18k cycles per tile); the TMA writes whole 128-byte lines instead.
            if (p.frag_stores && bn_valid == p.BN && !(p.ablate & 64)) {
                // ---- level 0 straight from registers: the 16x256b fragment gives four neighbouring threads one 32-byte run of
                // a row, so a warp-wide st.global.v2 writes eight whole sectors -- no staging in shared memory, no fence, and
                // no store box in the TMA unit, which then carries only the operand stages (write_probe: loads + these stores
                // 4.1 TB/s against 3.6 with the twenty boxes).  MEASURED SLOWER in the kernel (64 x 60x80, same box: 2565-2600 against
                // 2222-2227 us): the LSU already carries the level-1 and level-2 rows.  Experiment, flags bit26, parity-tested.
                for (int ci = half; ci < p.BN / 32; ci += EPI_SPLIT) {
#pragma unroll
                    for (int hr = 0; hr < 2; ++hr) {
                        uint32_t v[16];
                        tmem_ld_16x256b_x4(tmem_base + ((uint32_t)(32 * quarter + 16 * hr) << 16) + (uint32_t)acc * MAX_BN + (uint32_t)(32 * ci), v);
                        tmem_ld_wait();
                        const int r = mb * BM + 32 * quarter + 16 * hr + (lane >> 2);
                        float *dst = p.l0 + ((size_t)b * p.N + r) * p.N + j0 + 32 * ci + 2 * (lane & 3);
                        if (!(p.ablate & 1)) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 lo = make_float2(__uint_as_float(v[4 * j]) * scale, __uint_as_float(v[4 * j + 1]) * scale);
                                const float2 hi = make_float2(__uint_as_float(v[4 * j + 2]) * scale, __uint_as_float(v[4 * j + 3]) * scale);
                                float2 *d0 = reinterpret_cast<float2 *>(dst + 8 * j), *d1 = reinterpret_cast<float2 *>(dst + (size_t)8 * p.N + 8 * j);
                                if (r < p.N) { if (p.stream_l0) __stcs(d0, lo); else *d0 = lo; }
                                if (r + 8 < p.N) { if (p.stream_l0) __stcs(d1, hi); else *d1 = hi; }
                            }
                        }
                    }
                }
            } else if (p.qbox && !(p.ablate & 64)) {
                // ---- level 0, one box per lane quarter: the SM's TMA unit carries the operand stages and the volume's stores
                // and is ~90 % busy (profiles/r02/write_probe.txt, corr_trace_f16_8x60x80.txt); twenty {32 x 32} boxes per tile cost
                // it 2.4 us, four {32 x 32 x 5} boxes 1.5 us.  The two warps of the quarter stage their chunks (same swizzled
                // 4 KB images as below) side by side, meet, and one lane stores all of them.
                uint8_t *qb = epi + quarter * (p.BN / 32) * EPI_BUF_BYTES;
                const bool issuer = half == 0 && lane == 0;
                if (issuer) tma_store_wait_read<0>();      // the previous tile's box has left the buffer
                quarter_barrier(quarter);
                // chunks [0, qbox_h0) to the first warp, the rest to the second (which also takes the odd pooling strip)
                for (int ci = half ? p.qbox_h0 : 0; ci < (half ? p.BN / 32 : p.qbox_h0); ++ci) {
                    uint32_t v[32];
                    tmem_ld32(taddr + 32 * ci, v);
                    tmem_ld_wait();
                    uint8_t *buf = qb + ci * EPI_BUF_BYTES;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 o = make_float4(__uint_as_float(v[4 * q]) * scale, __uint_as_float(v[4 * q + 1]) * scale,
                                                     __uint_as_float(v[4 * q + 2]) * scale, __uint_as_float(v[4 * q + 3]) * scale);
                        *reinterpret_cast<float4 *>(buf + lane * 128 + ((q ^ (lane & 7)) << 4)) = o;
                    }
                }
                fence_async_smem();
                quarter_barrier(quarter);
                if (issuer) {
                    if (mb * BM + 32 * quarter < p.N && !(p.ablate & 1))
                        tma_store_3d(&tmap_c, qb, 0, b * p.N + mb * BM + 32 * quarter, j0 >> 5);
                    tma_store_commit();
                }
            } else if (bn_valid >= 32 && !(p.ablate & 64)) {
                const int nchunks = (bn_valid + 31) / 32;
                for (int ci = half; ci < nchunks; ci += EPI_SPLIT) {
                    const int c0 = min(ci * 32, bn_valid - 32);  // last chunk overlaps its neighbour (same values)
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    tmem_ld_wait();
                    uint8_t *buf = my_epi + (chunk_count & 1) * EPI_BUF_BYTES;
                    if (p.lsu_stores) {
                        // transpose through the (swizzled) buffer, then whole 128-byte rows with plain vector stores:
                        // lane <-> (row = 4*it + lane/8, 16-byte piece = lane%8), four rows per instruction.  These go
                        // through the LSU, not the TMA unit, so the operand loads never queue behind the volume's
                        // HBM-paced write stream (with TMA bulk stores the tile period was store time + MMA time)
                        __syncwarp();   // the previous chunk's reads of this buffer (two chunks ago) are done
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 o = make_float4(__uint_as_float(v[4 * q]) * scale, __uint_as_float(v[4 * q + 1]) * scale,
                                                         __uint_as_float(v[4 * q + 2]) * scale, __uint_as_float(v[4 * q + 3]) * scale);
                            *reinterpret_cast<float4 *>(buf + lane * 128 + ((q ^ (lane & 7)) << 4)) = o;
                        }
                        __syncwarp();
                        const int piece = lane & 7;
                        float *dst0 = p.l0 + ((size_t)b * p.N + (size_t)(mb * BM + 32 * quarter)) * p.N + j0 + c0 + 4 * piece;
                        float4 rowv[8];
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int r = 4 * it + (lane >> 3);
                            rowv[it] = *reinterpret_cast<const float4 *>(buf + r * 128 + ((piece ^ (r & 7)) << 4));
                        }
                        if (!(p.ablate & 1)) {
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const int r = 4 * it + (lane >> 3);
                                if (mb * BM + 32 * quarter + r < p.N) {
                                    float4 *d = reinterpret_cast<float4 *>(dst0 + (size_t)r * p.N);
                                    if (p.stream_l0) __stcs(d, rowv[it]); else *d = rowv[it];
                                }
                            }
                        }
                        ++chunk_count;
                        continue;
                    }
                    if (chunk_count >= 2) {  // the store issued two chunks ago has finished reading this buffer
                        if (lane == 0) tma_store_wait_read<1>();
                        __syncwarp();
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 o = make_float4(__uint_as_float(v[4 * q]) * scale, __uint_as_float(v[4 * q + 1]) * scale,
                                                     __uint_as_float(v[4 * q + 2]) * scale, __uint_as_float(v[4 * q + 3]) * scale);
                        *reinterpret_cast<float4 *>(buf + lane * 128 + ((q ^ (lane & 7)) << 4)) = o;
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0 && mb * BM + 32 * quarter < p.N && !(p.ablate & 1)) {  // (rows below the volume: nothing to store)
                        tma_store_3d(&tmap_c, buf, j0 + c0, mb * BM + 32 * quarter, b);
                        tma_store_commit();
                    }
                    ++chunk_count;
                }
            } else if (half == 0) {
                float *l0row = p.l0 + ((size_t)b * p.N + (row_ok ? i : 0)) * p.N;
                uint32_t v[32];
                tmem_ld32(taddr, v);
                tmem_ld_wait();
                if (row_ok && bn_valid > 0) store_row_chunk(l0row + j0, v, scale, bn_valid, vec4_l0 && (bn_valid % 4 == 0), false);
            }
            if (threadIdx.x == 0) TC_TRACE(tile_no, 13);
            // ---- level 1: 2x2 means straight from the accumulator rows
            if (p.R != 0 && !(p.ablate & 32)) {
                const int y0 = nb * p.R;
                float *l1map = p.l1 + ((size_t)b * p.N + (row_ok ? i : 0)) * p.h1 * p.w1;
                float prev1[16], prev2[8];   // deep fusion: the previous level-1 / level-2 row of this query
                int item = EPI_SPLIT - 1;  // an odd chunk count leaves the last warp of a quarter one item short: it starts here
                for (int pr = 0; pr < p.R; pr += 2) {
                    const int y = y0 + pr;
                    if (y + 1 >= p.h) break;  // warp-uniform
                    for (int xc = 0; xc < p.w; xc += 32) {
                        if (((item++) % EPI_SPLIT) != half) continue;
                        uint32_t ra[32], rc[32];
                        tmem_ld32(taddr + pr * p.w + xc, ra);
                        tmem_ld32(taddr + (pr + 1) * p.w + xc, rc);
                        tmem_ld_wait();
                        // ATen avg_pool2d: ((a + b) + c) + d, then / 4, on the stored (scaled) values
                        float o[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            float s = __uint_as_float(ra[2 * q]) * scale + __uint_as_float(ra[2 * q + 1]) * scale;
                            s += __uint_as_float(rc[2 * q]) * scale;
                            s += __uint_as_float(rc[2 * q + 1]) * scale;
                            o[q] = s * 0.25f;
                        }
                        if (p.l1_staged && p.w1 - (xc >> 1) >= 8) {   // (a narrow last strip is not worth the round trip)
                            // the lane's 16 values (64 B of ITS row) -> shared memory -> whole 64-byte runs, 8 rows per
                            // store instruction (lane <-> row 8*it + lane/4, piece lane%4) instead of 32 rows x 16 B
                            const int npool = min(16, p.w1 - (xc >> 1));
                            float *stg = p.l1_scratch ? reinterpret_cast<float *>(epi + p.l1_scratch + warp * 2048)
                                                      : reinterpret_cast<float *>(my_epi);
                            if (!p.l1_scratch && !p.lsu_stores && lane == 0) tma_store_wait_read<0>();   // the level-0 boxes have left the buffers
                            __syncwarp();
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                *reinterpret_cast<float4 *>(stg + lane * 16 + 4 * ((q + (lane >> 1)) & 3)) =
                                    make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
                            __syncwarp();
                            const int row0 = mb * BM + 32 * quarter;
                            if (vec4_l1) {
                                const int pc = lane & 3;
#pragma unroll
                                for (int it = 0; it < 4; ++it) {
                                    const int rr = 8 * it + (lane >> 2);
                                    const float4 val = *reinterpret_cast<const float4 *>(stg + rr * 16 + 4 * ((pc + (rr >> 1)) & 3));
                                    if (row0 + rr < p.N && 4 * pc < npool && !(p.ablate & 1))
                                        *reinterpret_cast<float4 *>(p.l1 + ((size_t)b * p.N + row0 + rr) * p.h1 * p.w1 +
                                                                    (size_t)(y >> 1) * p.w1 + (xc >> 1) + 4 * pc) = val;
                                }
                            } else {   // w1 even but not a multiple of 4: rows are only 8-byte aligned -> float2 pieces, 4 rows per store
                                const int pc = lane & 7;
#pragma unroll
                                for (int it = 0; it < 8; ++it) {
                                    const int rr = 4 * it + (lane >> 3);
                                    const float2 val = *reinterpret_cast<const float2 *>(
                                        stg + rr * 16 + 4 * (((pc >> 1) + (rr >> 1)) & 3) + 2 * (pc & 1));
                                    if (row0 + rr < p.N && 2 * pc < npool && !(p.ablate & 1))
                                        *reinterpret_cast<float2 *>(p.l1 + ((size_t)b * p.N + row0 + rr) * p.h1 * p.w1 +
                                                                    (size_t)(y >> 1) * p.w1 + (xc >> 1) + 2 * pc) = val;
                                }
                            }
                            __syncwarp();
                        } else if (row_ok && !(p.ablate & 1)) {
                            float *dst = l1map + (size_t)(y >> 1) * p.w1 + (xc >> 1);
                            const int npool = min(16, p.w1 - (xc >> 1));
                            if (vec4_l1) {
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    if (4 * q < npool)
                                        *(reinterpret_cast<float4 *>(dst) + q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
                            } else {
#pragma unroll
                                for (int q = 0; q < 8; ++q)
                                    if (2 * q < npool) *(reinterpret_cast<float2 *>(dst) + q) = make_float2(o[2 * q], o[2 * q + 1]);
                            }
                        }
                        if (p.pair2) {
                            // level 2 across the tile pair: the first tile parks its level-1 strip in TMEM columns
                            // [BN_mma, BN_mma + w1) of stage 0's window (free: BN_mma + w1 <= 256), the second tile pools
                            // ATen-style ((a + b) + c) + d, / 4 and stores 8 level-2 values per query and strip
                            const uint32_t park = tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)p.BN_mma + (uint32_t)(xc >> 1);
                            if ((nb & 1) == 0) {
                                uint32_t ov[16];
#pragma unroll
                                for (int q = 0; q < 16; ++q) ov[q] = __float_as_uint(o[q]);
                                tmem_st16(park, ov);
                            } else {
                                uint32_t pv[16];
                                tmem_ld16(park, pv);
                                float l2v[8];
#pragma unroll
                                for (int q = 0; q < 8; ++q) {
                                    float s2 = __uint_as_float(pv[2 * q]) + __uint_as_float(pv[2 * q + 1]);
                                    s2 += o[2 * q];
                                    s2 += o[2 * q + 1];
                                    l2v[q] = s2 * 0.25f;
                                }
                                const int r2g = nb >> 1, n2 = min(8, p.w2 - (xc >> 2));
                                if (row_ok && r2g < p.h2 && !(p.ablate & 1)) {
                                    float *d2 = p.l2 + (((size_t)b * p.N + i) * p.h2 + r2g) * p.w2 + (xc >> 2);   // w2 % 4 == 0
#pragma unroll
                                    for (int q = 0; q < 2; ++q)
                                        if (4 * q < n2)
                                            *(reinterpret_cast<float4 *>(d2) + q) = make_float4(l2v[4 * q], l2v[4 * q + 1], l2v[4 * q + 2], l2v[4 * q + 3]);
                                }
                                if (p.quad3) {
                                    // level 3 across the quad: the second tile parks this level-2 strip behind the level-1
                                    // park (16 columns per strip there, 8 per strip here), the fourth pools 4 level-3 values per query and strip
                                    const uint32_t park2 = tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(p.BN_mma + 16 * ((p.w + 31) >> 5)) + (uint32_t)(xc >> 2);
                                    if ((nb & 3) == 1) {
                                        uint32_t ov[8];
#pragma unroll
                                        for (int q = 0; q < 8; ++q) ov[q] = __float_as_uint(l2v[q]);
                                        tmem_st8(park2, ov);
                                    } else {
                                        uint32_t qv[8];
                                        tmem_ld8(park2, qv);
                                        const int r3g = nb >> 2, n3 = min(4, p.w3 - (xc >> 3));
                                        if (row_ok && r3g < p.h3 && !(p.ablate & 1)) {
                                            float *d3 = p.l3 + (((size_t)b * p.N + i) * p.h3 + r3g) * p.w3 + (xc >> 3);
#pragma unroll
                                            for (int q = 0; q < 4; ++q) {
                                                float s3 = __uint_as_float(qv[2 * q]) + __uint_as_float(qv[2 * q + 1]);
                                                s3 += l2v[2 * q];
                                                s3 += l2v[2 * q + 1];
                                                if (q < n3) d3[q] = s3 * 0.25f;
                                            }
                                        }
                                    }
                                }
                            }
                        }
                        if (p.deep) {   // w <= 32: this is the only 32-column strip of the row pair
                            // ATen avg_pool2d again, on the level-1 (then level-2) values just stored
                            const int r1g = y >> 1;
                            if (r1g & 1) {
                                float l2v[8];
#pragma unroll
                                for (int q = 0; q < 8; ++q) {
                                    float s2 = prev1[2 * q] + prev1[2 * q + 1];
                                    s2 += o[2 * q];
                                    s2 += o[2 * q + 1];
                                    l2v[q] = s2 * 0.25f;
                                }
                                const int r2g = r1g >> 1;
                                if (row_ok && r2g < p.h2 && !(p.ablate & 1)) {
                                    float *d2 = p.l2 + (((size_t)b * p.N + i) * p.h2 + r2g) * p.w2;   // w2 is even
                                    if ((p.w2 & 3) == 0) {   // 16-byte aligned rows: w2 = 4 or 8
#pragma unroll
                                        for (int q = 0; q < 2; ++q)
                                            if (4 * q < p.w2)
                                                *(reinterpret_cast<float4 *>(d2) + q) = make_float4(l2v[4 * q], l2v[4 * q + 1], l2v[4 * q + 2], l2v[4 * q + 3]);
                                    } else {
#pragma unroll
                                        for (int q = 0; q < 4; ++q)
                                            if (2 * q < p.w2) *(reinterpret_cast<float2 *>(d2) + q) = make_float2(l2v[2 * q], l2v[2 * q + 1]);
                                    }
                                }
                                if (p.deep > 1) {
                                    if (r2g & 1) {
                                        const int r3g = r2g >> 1;
                                        float l3v[4];
#pragma unroll
                                        for (int q = 0; q < 4; ++q) {
                                            float s3 = prev2[2 * q] + prev2[2 * q + 1];
                                            s3 += l2v[2 * q];
                                            s3 += l2v[2 * q + 1];
                                            l3v[q] = s3 * 0.25f;
                                        }
                                        if (row_ok && r3g < p.h3 && !(p.ablate & 1)) {
                                            float *d3 = p.l3 + (((size_t)b * p.N + i) * p.h3 + r3g) * p.w3;
                                            if (p.w3 == 4 && (reinterpret_cast<uintptr_t>(p.l3) & 15u) == 0) {
                                                *reinterpret_cast<float4 *>(d3) = make_float4(l3v[0], l3v[1], l3v[2], l3v[3]);
                                            } else {
#pragma unroll
                                                for (int q = 0; q < 4; ++q)
                                                    if (q < p.w3) d3[q] = l3v[q];
                                            }
                                        }
                                    } else {
#pragma unroll
                                        for (int q = 0; q < 8; ++q) prev2[q] = l2v[q];
                                    }
                                }
                            } else {
#pragma unroll
                                for (int q = 0; q < 16; ++q) prev1[q] = o[q];
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            if (CL == 2) mbar_arrive_cluster(mapa_rank0(&tempty[acc]));
            else mbar_arrive(&tempty[acc]);
            if (tile_no == 0 && threadIdx.x == 0) stamp(21);
            if (threadIdx.x == 0) TC_TRACE(tile_no, 14);
            if (kblocks <= 2 && lane == 0 && warp > 0) TC_TRACE(tile_no, warp == 3 ? 15 : (warp < 3 ? 4 + warp : 3 + warp));   // (slots of K blocks 2.. are free: epilogue done, warps 1,2 -> 5,6; 4..7 -> 7..10; 3 -> 15)
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    if (warp < EPI_WARPS && lane == 0) tma_store_wait_all();  // bulk stores must drain before the CTA's smem goes away
    tc_fence_before();
    __syncthreads();
    __syncwarp();
    if (CL == 2) cluster_sync_all();  // the peer may still commit onto this CTA's barriers until its own loop ends
    if (warp == MMA_WARP) {
        tc_fence_after();
        if (CL == 2) tmem_dealloc_2sm(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
    if (threadIdx.x == 0) stamp(22);
}

// ---- measured and dropped: A-resident variant --------------------------------------------------
// Round 1 also built a variant that keeps the 128 x D slice of fmap1 resident in shared memory (128 KB)
// for a whole range of target tiles, streaming only fmap2 (operand traffic per 128x160 tile 160 KB
// instead of 288 KB).  It was SLOWER (8 x 60x80: 426 us against 327 us; 1 x 80x124: 248 against 198):
// with 128 KB pinned only ~64 KB are left for the fmap2 ring, and at ~1.6 us of loaded TMA latency
// 60 KB in flight sustain ~37 KB/us per SM -- the streaming kernel keeps 144 KB in flight.  The per-tile
// timeline (scripts/corr_trace.py at that commit) showed the main loop waiting on loads, the epilogue
// (3.4-3.9 us per tile) hidden behind it.  What did help both variants: ONE 3-D TMA instruction per
// operand and stage (atoms3d / b3d above) instead of 9-12 per-atom boxes -- a single thread issues a
// cp.async.bulk.tensor only about every 100 cycles, which was the main-loop bound (368 -> 327 us).
// Next lever (not done): cta_group::2 pairs (M = 256 across two SMs, each loading half of the fmap2 tile).

CF_DEFINE_TRACE_SETTER(cf_trace_buffer_corr)

// ---- fp16 operand copies ---------------------------------------------------------------------------------
// amax[t * B + b] = max |x| over batch item b of feature map t (t = 0, 1), as the bit pattern of a non-negative float
// (integer max == float max); zeroed by the host side before the launch.
__global__ void __launch_bounds__(256)
fmap_absmax_kernel(const float *__restrict__ f1, const float *__restrict__ f2, int64_t per_item, int B, int b0, int nbc,
                   unsigned *__restrict__ amax) {
    // this launch covers batch items [b0, b0 + nbc) of both maps: blockIdx.y = t * nbc + (b - b0)
    const int t = (int)blockIdx.y / nbc, b = b0 + (int)blockIdx.y % nbc;
    const int item = t * B + b;
    const float *src = (t == 0 ? f1 : f2) + (int64_t)b * per_item;
    float m = 0.f;
    const int64_t n4 = per_item >> 2;
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(s4 + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    m = warp_max(m);
    __shared__ float sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) m = fmaxf(m, sh[k]);
        if (m > 0.f) atomicMax(amax + item, __float_as_uint(fminf(m, 3.0e38f)));
    }
}
// x -> fp16(x * 2^shift), shift chosen per batch item so that the largest magnitude lands in [2^13, 2^14): every finite
// input fits fp16 (round to nearest: the same 11-bit significand a TF32 operand keeps), the scaling is exact, and
// inv_scale[item] = 2^-shift lets the GEMM epilogue undo it.  per_item % 4 == 0.
__global__ void __launch_bounds__(256)
fmap_to_half_kernel(const float *__restrict__ f1, const float *__restrict__ f2, int64_t per_item, int B, int b0, int nbc,
                    const unsigned *__restrict__ amax, __half *__restrict__ h1, __half *__restrict__ h2,
                    float *__restrict__ inv_scale) {
    const int t = (int)blockIdx.y / nbc, b = b0 + (int)blockIdx.y % nbc;
    const int item = t * B + b;
    const bool first = t == 0;
    const int64_t base = (int64_t)b * per_item;
    const float *src = (first ? f1 : f2) + base;
    __half *dst = (first ? h1 : h2) + base;
    const float mx = __uint_as_float(__ldg(amax + item));
    const int shift = mx > 0.f ? 13 - ilogbf(mx) : 0;
    const float up = ldexpf(1.f, shift > 126 ? 126 : shift);      // (a tiny maximum: clamp the exponent, still exact)
    if (blockIdx.x == 0 && threadIdx.x == 0) inv_scale[item] = 1.f / up;
    const int64_t n4 = per_item >> 2;
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    uint2 *d4 = reinterpret_cast<uint2 *>(dst);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(s4 + i);
        const __half2 lo = __floats2half2_rn(v.x * up, v.y * up), hi = __floats2half2_rn(v.z * up, v.w * up);
        uint2 o;
        o.x = *reinterpret_cast<const unsigned *>(&lo);
        o.y = *reinterpret_cast<const unsigned *>(&hi);
        d4[i] = o;
    }
}

// One pass instead of two (round 2): a CTA keeps its 32 KB slice of a feature map in REGISTERS (8 float4 per thread),
// publishes the slice's maximum (atomicMax on the item's slot + an arrival counter), waits until all slices of the item have
// arrived, then scales and narrows out of registers -- the fp32 maps cross the HBM interface once.  The two-kernel form
// (maximum pass, then conversion pass, chunked so that the second read hit the L2) measured 342 us for 64 x 2 maps of 256 x
// 60x80 under ncu -- 15 % of the whole pyramid build.  Units are ordered by item and there are no more slices per item
// than CTAs, so a CTA never waits for a slice that is queued behind its own (all CTAs of the launch fit on the chip at once).
constexpr int CV_THREADS = 256, CV_F4 = 8;
// (round 2, later: the loads of a CTA's NEXT slice are issued before it waits for the current item's arrival counter --
//  without that all resident CTAs loaded, waited and stored in lockstep waves, reads and writes taking turns on the bus)
__global__ void __launch_bounds__(CV_THREADS, 3)
fmap_to_half_fused_kernel(const float *__restrict__ f1, const float *__restrict__ f2, int64_t per_item, int B, int parts,
                          unsigned *amax, int *arrived, __half *__restrict__ h1, __half *__restrict__ h2,
                          float *__restrict__ inv_scale) {
    __shared__ float s_red[2][CV_THREADS / 32];
    __shared__ float s_up[2];
    const int64_t units = (int64_t)2 * B * parts, n4 = per_item >> 2;
    struct Unit {
        const float4 *s4;
        uint2 *d4;
        int64_t i0;
        int item, part;
    };
    auto unit_of = [&](int64_t u) {
        Unit t;
        t.item = (int)(u / parts);
        t.part = (int)(u - (int64_t)t.item * parts);
        const bool first = t.item < B;
        const int64_t base = (int64_t)(first ? t.item : t.item - B) * per_item;
        t.s4 = reinterpret_cast<const float4 *>((first ? f1 : f2) + base);
        t.d4 = reinterpret_cast<uint2 *>((first ? h1 : h2) + base);
        t.i0 = (int64_t)t.part * (CV_THREADS * CV_F4) + threadIdx.x;
        return t;
    };
    auto load = [&](const Unit &t, float4 (&v)[CV_F4]) {
#pragma unroll
        for (int k = 0; k < CV_F4; ++k) {
            const int64_t i = t.i0 + (int64_t)k * CV_THREADS;
            v[k] = i < n4 ? __ldg(t.s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    // the slice's maximum -> the item's slot, arrival counted (thread 0 of the CTA); `slot` alternates between units
    auto publish = [&](const Unit &t, const float4 (&v)[CV_F4], int slot) {
        float m = 0.f;
#pragma unroll
        for (int k = 0; k < CV_F4; ++k)
            m = fmaxf(fmaxf(m, fmaxf(fabsf(v[k].x), fabsf(v[k].y))), fmaxf(fabsf(v[k].z), fabsf(v[k].w)));
        m = warp_max(m);
        if ((threadIdx.x & 31) == 0) s_red[slot][threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < CV_THREADS / 32; ++k) m = fmaxf(m, s_red[slot][k]);
            if (m > 0.f) atomicMax(amax + t.item, __float_as_uint(fminf(m, 3.0e38f)));
            __threadfence();
            atomicAdd(arrived + t.item, 1);
        }
    };
    // wait until every slice of the item has arrived, then scale and narrow out of registers
    auto finish = [&](const Unit &t, const float4 (&v)[CV_F4], int slot) {
        if (threadIdx.x == 0) {
            unsigned spins = 0;
            while (*reinterpret_cast<volatile int *>(arrived + t.item) < parts) {
                __nanosleep(64);
                if (++spins > (1u << 24)) __trap();
            }
            __threadfence();
            const float mx = __uint_as_float(*reinterpret_cast<volatile unsigned *>(amax + t.item));
            const int shift = mx > 0.f ? 13 - ilogbf(mx) : 0;
            const float up = ldexpf(1.f, shift > 126 ? 126 : shift);      // (a tiny maximum: clamp the exponent, still exact)
            s_up[slot] = up;
            if (t.part == 0) inv_scale[t.item] = 1.f / up;
        }
        __syncthreads();
        const float up = s_up[slot];
#pragma unroll
        for (int k = 0; k < CV_F4; ++k) {
            const int64_t i = t.i0 + (int64_t)k * CV_THREADS;
            if (i < n4) {
                const __half2 lo = __floats2half2_rn(v[k].x * up, v[k].y * up), hi = __floats2half2_rn(v[k].z * up, v[k].w * up);
                uint2 o;
                o.x = *reinterpret_cast<const unsigned *>(&lo);
                o.y = *reinterpret_cast<const unsigned *>(&hi);
                t.d4[i] = o;
            }
        }
    };
    // two register sets in turn: publish(cur) -> load(next) -> finish(cur)
    float4 va[CV_F4], vb[CV_F4];
    int64_t u = blockIdx.x;
    if (u >= units) return;
    Unit ta = unit_of(u), tb = ta;
    load(ta, va);
    for (;;) {
        publish(ta, va, 0);
        const bool more_b = u + gridDim.x < units;
        if (more_b) { tb = unit_of(u + gridDim.x); load(tb, vb); }
        finish(ta, va, 0);
        if (!more_b) break;
        u += gridDim.x;
        publish(tb, vb, 1);
        const bool more_a = u + gridDim.x < units;
        if (more_a) { ta = unit_of(u + gridDim.x); load(ta, va); }
        finish(tb, vb, 1);
        if (!more_a) break;
        u += gridDim.x;
    }
}

// ---- host side --------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D map over a feature map viewed as [B*D rows, N cols], box 32 x 32, 128B swizzle / 32B atoms
// (fp16 operand copies: 2-byte elements, boxes of 64 columns, plain 128B swizzle)
static int make_fmap_tmap(CUtensorMap *m, const void *base, int B, int D, int N, CUtensorMapDataType dt, int box_rows = 32) {
    EncodeTiledFn fn = encode_fn();
    CF_REQUIRE(fn, CF_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    const bool half = dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const size_t elt = half ? 2 : 4;
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)B * D};
    cuuint64_t strides[1] = {(cuuint64_t)N * elt};
    cuuint32_t box[2] = {(cuuint32_t)(128 / elt), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    half ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CF_REQUIRE(r == CUDA_SUCCESS, CF_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return CF_OK;
}

// The same feature map as a 3-D tensor {32 cols, B*D rows, N/32 column atoms} (N % 32 == 0): a box
// {32, rows, atoms} lands in shared memory as [atom][row][32 cols] -- the UMMA MN-major layout -- in ONE instruction
static int make_fmap_tmap3(CUtensorMap *m, const void *base, int B, int D, int N, CUtensorMapDataType dt, int box_rows,
                           int box_atoms) {
    EncodeTiledFn fn = encode_fn();
    CF_REQUIRE(fn, CF_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    const bool half = dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const size_t elt = half ? 2 : 4;
    const cuuint32_t bc = (cuuint32_t)(128 / elt);
    cuuint64_t dims[3] = {bc, (cuuint64_t)B * D, (cuuint64_t)N / bc};
    cuuint64_t strides[2] = {(cuuint64_t)N * elt, 128};
    cuuint32_t box[3] = {bc, (cuuint32_t)box_rows, (cuuint32_t)box_atoms};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, dt, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    half ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CF_REQUIRE(r == CUDA_SUCCESS, CF_ERR_CUDA, "cuTensorMapEncodeTiled (3-D feature map) failed with CUresult %d", (int)r);
    return CF_OK;
}

// the level-0 volume as {32 cols, B*N rows, N/32 column atoms} (N % 32 == 0): a box {32, 32, atoms} stores `atoms` swizzled
// 32 x 32 images that lie side by side in shared memory with ONE instruction
static int make_volume_tmap_atoms(CUtensorMap *m, float *base, int B, int N, int atoms) {
    EncodeTiledFn fn = encode_fn();
    CF_REQUIRE(fn, CF_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {32, (cuuint64_t)B * N, (cuuint64_t)N / 32};
    cuuint64_t strides[2] = {(cuuint64_t)N * sizeof(float), 128};
    cuuint32_t box[3] = {32, 32, (cuuint32_t)atoms};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CF_REQUIRE(r == CUDA_SUCCESS, CF_ERR_CUDA, "cuTensorMapEncodeTiled (volume, atoms) failed with CUresult %d", (int)r);
    return CF_OK;
}

// fp16 feature map as {32 cols, B*D rows, N/32 atoms of 64 bytes}, 64-byte swizzle (N % 32 == 0)
static int make_fmap_tmap3_sw64(CUtensorMap *m, const void *base, int B, int D, int N, int box_rows, int box_atoms) {
    EncodeTiledFn fn = encode_fn();
    CF_REQUIRE(fn, CF_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {32, (cuuint64_t)B * D, (cuuint64_t)N / 32};
    cuuint64_t strides[2] = {(cuuint64_t)N * 2, 64};
    cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)box_atoms};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CF_REQUIRE(r == CUDA_SUCCESS, CF_ERR_CUDA, "cuTensorMapEncodeTiled (64B-swizzled feature map) failed with CUresult %d", (int)r);
    return CF_OK;
}

// 3-D map over the level-0 volume [B][N][N], box {32 cols, 32 rows, 1}, 128B swizzle (store side)
static int make_volume_tmap(CUtensorMap *m, float *base, int B, int N) {
    EncodeTiledFn fn = encode_fn();
    CF_REQUIRE(fn, CF_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)N * sizeof(float), (cuuint64_t)N * N * sizeof(float)};
    cuuint32_t box[3] = {32, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CF_REQUIRE(r == CUDA_SUCCESS, CF_ERR_CUDA, "cuTensorMapEncodeTiled (volume) failed with CUresult %d", (int)r);
    return CF_OK;
}

}  // namespace tc

bool corr_tensor_core_supported(int D, int h, int w) {
    return D % tc::BK == 0 && ((int64_t)h * w) % 4 == 0;
}

// fp16 operand copies of both feature maps + per-item maxima and inverse scales (CF_CORR_F16 / CF_CORR_AUTO)
size_t corr_tc_workspace_bytes(int B, int D, int h, int w) {
    const size_t per_map = align_up((size_t)B * D * h * w * sizeof(__half), 256);
    return 2 * per_map + align_up((size_t)6 * B * sizeof(float), 256);   // maxima, inverse scales, arrival counters
}

// flags (debug/experiments, env CF_TC_FLAGS): bit1 = encode the tensor maps as plain FLOAT32 (operands are
// then truncated, not rounded, to TF32), bit2 = never fuse the pooling, bit3 = fuse level 1 only, bit4 = per-atom 2-D TMA boxes even when N % 32 == 0.
int corr_volume_tensor_core(const float *f1, const float *f2, int B, int D, int h, int w, float scale, float *level0,
                            float *level1, float *level2, float *level3, int precision, void *ws, size_t ws_bytes, int flags,
                            int *fused_levels, cudaStream_t stream) {
    using namespace tc;
    CF_REQUIRE(precision == CF_CORR_TF32 || precision == CF_CORR_F16 || precision == CF_CORR_AUTO, CF_ERR_UNSUPPORTED,
               "cf_corr_build: CF_CORR_3XTF32 is not implemented yet");
    CF_REQUIRE(aligned16(f1) && aligned16(f2) && aligned16(level0), CF_ERR_ALIGN, "cf_corr_build: tensors must be 16-byte aligned");
    const int N = h * w;
    // AUTO resolves to TF32: the fp16-operand variant halves the MMA time and the operand bytes and agrees with
    // TF32 to summation-order noise (same 11-bit operand significands), but measured NO faster on the B200
    // (8 x 60x80: GEMM 283 against 277 us, plus 32 us of conversion) -- the kernel is bound on the volume's write side
    // AUTO (round 2): fp16 operands where they measured faster on the B200 -- wide maps whose N is a whole number of
    // 64-column fp16 boxes and enough tiles to amortise the conversion pass (8 x 60x80: 288 against 319 us, 64 x 60x80: 2375
    // against 2580 us) -- TF32 elsewhere (64 x 24x32: 61 against 108 us, 64 x 36x44: 336 against 374, 1 x 80x124: 183 against 205)
    if (precision == CF_CORR_AUTO)
        precision = (N % 64 == 0 && N >= 2048 && (int64_t)B * N >= 19200 && D % BK_F16 == 0 && ws != nullptr &&
                     ws_bytes >= corr_tc_workspace_bytes(B, D, h, w)) ? CF_CORR_F16 : CF_CORR_TF32;
    if (precision == CF_CORR_F16 && D % BK_F16 != 0) precision = CF_CORR_TF32;   // the fp16 stages hold 128 K rows
    const bool f16 = precision == CF_CORR_F16;
    const int BC = f16 ? 64 : 32;   // operand columns per TMA box
    const void *a = f1, *bm = f2;
    const float *inv_scale = nullptr;
    if (f16) {
        const size_t need = corr_tc_workspace_bytes(B, D, h, w);
        CF_REQUIRE(ws && ws_bytes >= need, CF_ERR_WORKSPACE, "cf_corr_build: workspace too small for CF_CORR_F16 (%zu < %zu)",
                   ws_bytes, need);
        CF_REQUIRE(aligned16(ws), CF_ERR_ALIGN, "cf_corr_build: workspace not 16-byte aligned");
        const int64_t per_item = (int64_t)D * N;   // N % 4 == 0 (corr_tensor_core_supported)
        const size_t per_map = align_up((size_t)B * per_item * sizeof(__half), 256);
        __half *h1 = reinterpret_cast<__half *>(ws), *h2 = reinterpret_cast<__half *>(reinterpret_cast<char *>(ws) + per_map);
        unsigned *amax = reinterpret_cast<unsigned *>(reinterpret_cast<char *>(ws) + 2 * per_map);
        float *inv = reinterpret_cast<float *>(amax + 2 * B);
        int *arrived = reinterpret_cast<int *>(inv + 2 * B);
        const int parts = (int)ceil_div(per_item / 4, CV_THREADS * CV_F4);
        int resident = 0;
        CF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, fmap_to_half_fused_kernel, CV_THREADS, 0));
        const int64_t cap = (int64_t)resident * sm_count();
        if (parts <= cap && !(flags & (1 << 16))) {   // (flags bit16: the two-pass form)
            // one pass: every slice of an item is in flight at once (parts <= CTAs of the launch)
            CF_CUDA(cudaMemsetAsync(amax, 0, sizeof(unsigned) * 2 * B, stream));
            CF_CUDA(cudaMemsetAsync(arrived, 0, sizeof(int) * 2 * B, stream));
            const int64_t units = (int64_t)2 * B * parts;
            fmap_to_half_fused_kernel<<<(unsigned)(units < cap ? units : cap), CV_THREADS, 0, stream>>>(f1, f2, per_item, B, parts, amax,
                                                                                                   arrived, h1, h2, inv);
            CF_LAUNCH_CHECK("fmap_to_half_fused_kernel");
        } else {
        CF_CUDA(cudaMemsetAsync(amax, 0, sizeof(unsigned) * 2 * B, stream));
        // groups of batch items whose fp32 maps fit the L2 together (<= 48 MB): the conversion pass re-reads what the
        // maximum pass just read out of L2, so the feature maps cross the HBM interface once, not twice
        int nbc = (int)((48ll << 20) / (2 * per_item * (int64_t)sizeof(float)));
        if (nbc < 1) nbc = 1;
        if (nbc > B) nbc = B;
        for (int b0 = 0; b0 < B; b0 += nbc) {
            const int nb_ = B - b0 < nbc ? B - b0 : nbc;
            int64_t bx = ceil_div(per_item / 4, 256 * 8);
            const int64_t cap = ceil_div(8 * (int64_t)sm_count(), 2 * nb_);
            if (bx > cap) bx = cap;
            if (bx < 1) bx = 1;
            dim3 grid((unsigned)bx, (unsigned)(2 * nb_));
            fmap_absmax_kernel<<<grid, 256, 0, stream>>>(f1, f2, per_item, B, b0, nb_, amax);
            CF_LAUNCH_CHECK("fmap_absmax_kernel");
            fmap_to_half_kernel<<<grid, 256, 0, stream>>>(f1, f2, per_item, B, b0, nb_, amax, h1, h2, inv);
            CF_LAUNCH_CHECK("fmap_to_half_kernel");
        }
        }
        a = h1; bm = h2; inv_scale = inv;
    }
    Params p{};
    p.B = B; p.D = D; p.N = N; p.h = h; p.w = w; p.scale = scale; p.l0 = level0; p.l1 = level1;
    p.inv_scale = inv_scale;
    p.h1 = h / 2; p.w1 = w / 2;
    // R = number of whole target rows per tile (even).  The epilogue reads 32-column chunks, so the
    // last chunk of the last row must stay inside the 256-column accumulator stage.
    int R = 0;
    if (level1 != nullptr && !(flags & 4) && (w % 4 == 0) && (h % 2 == 0) && 2 * w <= MAX_BN) {
        R = 2 * (MAX_BN / (2 * w));
        if (R > h) R = h;
        while (R > 0 && (R - 1) * w + (int)align_up(w, 32) > MAX_BN) R -= 2;
        // narrow maps: whole groups of 8 (or 4) rows per tile, so that levels 2 and 3 can be pooled in the epilogue too
        if (!(flags & 8) && w <= 32 && w % 8 == 0) {
            if (R >= 8 && h % 8 == 0) R &= ~7;
            else if (R >= 4 && h % 4 == 0) R &= ~3;
        }
    }
    // levels this launch produces beyond level 0.  Levels 2 / 3 ride along when a tile holds whole groups of 4 / 8
    // target rows (R % 4 == 0 / R % 8 == 0: maps up to 64 / 32 wide) and the widths halve evenly (flags bit3: off)
    p.deep = 0;
    if (R > 0 && !(flags & 8) && w <= 32 && level2 != nullptr && aligned16(level2) && R % 4 == 0 && w % 8 == 0 && h % 4 == 0) {
        p.deep = 1;
        if (level3 != nullptr && R % 8 == 0 && w % 8 == 0 && h % 8 == 0) p.deep = 2;
    }
    p.h2 = h / 4; p.w2 = w / 4; p.h3 = h / 8; p.w3 = w / 8;
    p.l2 = level2; p.l3 = level3;
    // wide maps (R == 2): level 2 out of pairs of tiles (kernel: pair2).  Needs an even number of target-row pairs, level-2
    // rows that start 16-byte aligned per 32-column strip, and w1 free TMEM columns behind stage 0's accumulator
    p.pair2 = 0;
    if (R == 2 && p.deep == 0 && !(flags & 8) && !(flags & 64) && level2 != nullptr && aligned16(level2) && h % 4 == 0 && w % 16 == 0 &&
        (int)align_up(R * w, 16) + w / 2 <= MAX_BN)
        p.pair2 = 1;
    // ... and level 3 out of quads of tiles (kernel: quad3): needs level-2 rows that pool inside a 32-column strip (w % 16 == 0
    // gives 4 or 8 level-2 values per strip) and w2 more free TMEM columns; h / 8 level-3 rows (a last half quad only parks)
    // MEASURED SLOWER (64 x 60x80: 2571 against 2007-2026 us, 8 x 60x80: 293 against 255): it saves the 89 us pooling kernel, but a
    // lane's four level-3 values are 16 bytes of a 280-byte per-query map -- 32 scattered partial-sector writes per store
    // instruction where the pooling kernel writes level 3 as one dense stream.  Parity-tested, on request only (flags bit25).
    p.quad3 = 0;
    if (p.pair2 && level3 != nullptr && (flags & (1 << 25)) && h >= 8 && w >= 8 &&
        (int)align_up(R * w, 16) + 24 * (int)ceil_div(w, 32) <= MAX_BN)
        p.quad3 = 1;
    *fused_levels = R > 0 ? 1 + (p.quad3 ? 2 : (p.pair2 ? 1 : p.deep)) : 0;
    p.stream_l0 = (int64_t)B * N * N * 4 > (64ll << 20);
    if (R > 0) {
        p.R = R;
        p.BN = R * w;
        p.tiles_n = (int)ceil_div(h, R);
    } else {
        p.R = 0;
        p.BN = N < MAX_BN ? (int)align_up(N, 16) : MAX_BN;
        p.tiles_n = (int)ceil_div(N, p.BN);
    }
    p.BN_mma = (int)align_up(p.BN, 16);
    p.n_boxes_b = (int)ceil_div(p.BN_mma, BC);
    p.tiles_m = (int)ceil_div(N, BM);
    p.tiles_nv = p.quad3 ? (int)align_up(p.tiles_n, 4) : p.tiles_n;
    const int64_t total = (int64_t)B * p.tiles_m * p.tiles_nv;
    CF_REQUIRE(total < (1ll << 31), CF_ERR_INVALID_ARG, "cf_corr_build: too many tiles");
    p.total_tiles = (int)total;

    const CUtensorMapDataType dt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                       : ((flags & 2) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32);
    p.atoms3d = (N % BC == 0) && !(flags & 16);
    p.b3d = p.atoms3d && (p.BN % BC == 0 || p.tiles_n == 1);
    const bool pair_req = (flags & 64) && !f16;
    const int bk = f16 ? BK_F16 : ((pair_req || D % BK_TF32 != 0) ? BK : BK_TF32);          // K rows per stage
    const int boxb = 128 * bk;                 // bytes per TMA box
    const int a_bytes = (BM / BC) * boxb;
    p.stage_bytes = a_bytes + p.n_boxes_b * boxb;
    p.b_sw64 = 0;
    // (the 128-byte-swizzled slice is one 3-D box only when tiles start on 64-column boundaries -- b3d; with 160-column tiles
    //  it took three 2-D boxes per stage)
    if (f16 && p.atoms3d && N % 32 == 0 && !pair_req && !(flags & (1 << 17)) && p.BN_mma % 64 == 32 && p.BN % 32 == 0) {
        p.b_sw64 = 1;
        p.stage_bytes = a_bytes + (p.BN_mma / 32) * bk * 64;
    }
    // two epilogue warps per TMEM lane quarter when the ring still gets >= 5 stages beside their 64 KB of buffers (fp16
    // operands); the deep pooling keeps per-thread row state and stays on one warp per quarter (flags bit7: force 1)
    // (measured slower again, also with the half-size fp16 stages: 396 against 359 us at 8 x 60x80 -- only on request)
    // round 2: with 64-row fp16 stages the operand loads take 4 x 0.45 us and the MMAs 1.9 us per 128x160 tile; the
    // epilogue on four warps (3.6 us per tile, corr_trace) became the critical path -> two warps per lane quarter by
    // default for fp16 operands (flags bit7 now forces one)
    const int es = (f16 && p.deep == 0 && !(flags & 128) && (SMEM_LIMIT - smem_fixed(2)) / p.stage_bytes >= 2) ? 2 : 1;
    p.stages = (SMEM_LIMIT - smem_fixed(es)) / p.stage_bytes;
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    // pairs of query tiles as one 256-row UMMA across two CTAs: measured slower than one CTA per tile on the B200
    // (8 x 60x80: 359 against 322 us, see the kernel's header), so only on request (flags bit6)
    const bool pair = (flags & 64) && p.tiles_m >= 2 && !f16;
    p.ablate = (flags >> 8) & 255;
    // level-0 rows through the LSU instead of TMA bulk stores: helps where the operand loads already need many TMA
    // instructions per stage (N % 32 != 0: 64 x 36x44 398 -> 367 us), costs 2-5 % elsewhere (flags bit0: always TMA)
    p.lsu_stores = (!(flags & 1) && !p.atoms3d && N % 4 == 0) ? 1 : 0;
    // round 2, measured and NOT adopted for fp16 operands: (a) level-0 rows through the LSU (flags bit5): 64 x 60x80 2617
    // against 2515 us; (b) one bulk store per 32-column chunk for all 128 rows of the tile (box {32, 128}, the four
    // epilogue warps of a lane-quarter group meeting at a named barrier; 5 store instructions per tile instead of 20):
    // 2704 against 2375 us (TF32: 2744 against 2580) -- the stores' instruction count is not what holds the epilogue up
    if (f16 && N % 4 == 0 && (flags & 32)) p.lsu_stores = 1;
    // (d) what the data movement alone costs (scripts/experiments/write_probe.cu -> profiles/r02/write_probe.txt, no MMAs, no
    // epilogue arithmetic, per 128x160 tile and SM): the two 80 KB operand stages 1.12 us; the 20 level-0 boxes 2.4-2.5 us
    // (4.9-5.0 TB/s chip-wide, the same with 2 CTAs per SM or 2-atom boxes: ~3.75 ns per 128-byte box row); both together
    // 3.27 us -- the TMA unit mostly serialises them, and the kernel's 3.75 us per tile (with levels 1-2 on top) is within 15 %
    // of that.  A write-only stream reaches 6.1-6.2 TB/s (memset 7.0), one bulk store per whole 640-byte tile row 5.9 TB/s,
    // loads + row stores 2.5 us per tile.  Two epilogues built on row stores (padded row-major staging, 16 rows per lane
    // quarter single-buffered / 8 rows double-buffered, the warp pair of a quarter meeting at a named barrier, lanes issuing
    // their rows) passed the parity tests and were SLOWER: 2685 / 3350 against 2100-2190 us at 64 x 60x80 -- beside two 80 KB
    // operand stages only 64 KB are left, a lane quarter's rows (20 KB) cannot all be staged at once, and the extra passes
    // (barriers, TMEM re-reads at the register cap of 168) lengthen the epilogue's serial chain more than the cheaper stores
    // shorten the TMA queue.  (e) level-0 chunks split between TMA (half-0 warps) and LSU (half-1 warps) or the reverse:
    // 2280 / 2318 against 2190 us.  (f) other splits of the 5 chunks between the two warps of a lane quarter than the
    // alternating 3 | 2 (the second warp also takes 2 of the 3 pooling strips): 4 | 1 2256-2276, 2 | 3 2240, 5 | 0 2372 against
    // 2197 us.  All removed again.  What the per-tile timeline shows (scripts/corr_trace.py with the stamps of four epilogue
    // warps, profiles/r02/corr_trace_f16_8x60x80.txt): an epilogue warp is busy 2.75-3.0 us per tile and then WAITS 0.6-0.9 us for
    // the next accumulator; the tile period (3.6-3.9 us) is the chain  accumulator drained by all 8 warps -> operand stages
    // of the tile after next land (the ring of two 80 KB stages holds exactly one tile's K, and these loads queue in the TMA
    // unit behind the store boxes) -> its MMAs run (~2 us, bound by the operand fetch from shared memory) -> epilogue.
    // (c) one epilogue pass per 32-column strip of an R == 2 tile -- both level-0 boxes, level 1 and level 2 from ONE pair of
    // TMEM loads instead of 5 chunk loads + 6 strip loads per warp, no drain between the phases: SLOWER on the same box
    // (64 x 60x80: fp16 2668 against 2558 us, TF32 2482 against 2392 us) -- each warp then has a single pair of staging
    // buffers in flight and waits for its bulk stores once per strip.  Removed again; the two-phase epilogue stays.
    // level-1 rows through shared memory into 64-byte runs (flags bit13: per-lane 16-byte stores): 64 x 24x32
    // 66.7 -> 60.8 us, 8 x 60x80 314 -> 309 us
    p.l1_staged = (p.R > 0 && p.w1 % 2 == 0 && !(flags & 8192) && !(flags & (1 << 24))) ? 1 : 0;   // (bit24: off, without bit13's ablation side effect)
    // level 0 as one multi-atom box per lane quarter (kernel: qbox).  Needs whole tiles (h % R == 0), 32-column atoms that
    // start on atom boundaries, whole lane quarters (N % 32 == 0) and the tile's level 0 (BN x 128 x 4 B) inside the shared
    // memory left beside the operand ring -- for 160 columns that takes the 64-byte-swizzled fmap2 slice (b_sw64).  The
    // level-1 scratch has no room then: level-1 rows leave as per-lane 16-byte stores.  Measured SLOWER at 64 x 60x80: 2476
    // against 1985 us for the twenty single boxes on the same box (one buffer per lane quarter: every tile waits for its
    // predecessor's box to leave, and two named barriers) -- only on request (flags bit19).  flags bit17 switches b_sw64 off.
    // (Also measured: the epilogue's shared-memory accesses as explicit st.shared / ld.shared asm instead of the generic ST.E / LD.E
    //  the compiler emits for the aligned-up buffer pointer: 2545-2568 against 2221-2243 us on one box -- volatile asm with a memory
    //  clobber pins every store in program order between the TMEM loads and fences; the plain C++ dereferences stay.)
    // Most of that loss is the level-1 rows: per-lane 16-byte stores alone (flags bit24, single boxes) cost 2485 against 2015 us
    // at this shape.  A variant with one {32 x 32 x 3 | 2} box per WARP (no barrier, 8 stores per tile; write_probe: 5.8 against
    // 4.9 TB/s for the stores alone, 4.3 against 3.7 with the operand loads) measured 2630 us with the same level-1 stores,
    // i.e. still behind the single boxes -- removed again.
    p.frag_stores = (R > 0 && (flags & (1 << 26)) && h % R == 0 && N % 8 == 0 && p.BN % 32 == 0 && !pair) ? 1 : 0;
    p.qbox = 0;
    if (es == 2 && R > 0 && (flags & (1 << 19)) && h % R == 0 && N % 32 == 0 && p.BN % 32 == 0 && !p.lsu_stores && !pair &&
        p.stages * p.stage_bytes + 4 * (p.BN / 32) * EPI_BUF_BYTES + 1024 + 256 <= SMEM_LIMIT) {
        p.qbox = 1;
        p.l1_staged = 0;
        const int nch = p.BN / 32, ns = (R / 2) * (int)ceil_div(w, 32);
        // a pooling strip costs a warp about twice a chunk (timeline): the second warp gets one strip more when their count is odd
        p.qbox_h0 = (nch + 1) / 2 + ((ns & 1) ? 1 : 0);
        if (p.qbox_h0 > nch) p.qbox_h0 = nch;
        if ((flags >> 20) & 15) p.qbox_h0 = min(nch, ((flags >> 20) & 15) - 1);
    }
    p.tiles_mp = (int)ceil_div(p.tiles_m, 2);
    p.b_half = (int)ceil_div(p.BN_mma / 2, 32);
    if (pair) {
        p.total_tiles = B * p.tiles_mp * p.tiles_n;
        p.b3d = p.b3d && (p.BN_mma / 2) % 32 == 0;
        p.stage_bytes = A_BYTES + p.b_half * BOX_BYTES;
        p.stages = (SMEM_LIMIT - smem_fixed(1)) / p.stage_bytes;
        if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    }
    p.epi_bytes = p.qbox ? 4 * (p.BN / 32) * EPI_BUF_BYTES : epi_bytes(es);
    p.l1_scratch = 0;
    if (es == 2 && !p.qbox && p.l1_staged && p.stages * p.stage_bytes + epi_bytes(2) + 8 * 2048 + 1024 + 256 <= SMEM_LIMIT) {
        p.l1_scratch = epi_bytes(2);
        p.epi_bytes += 8 * 2048;
    }
    const int smem_bytes = p.stages * p.stage_bytes + p.epi_bytes + 1024 + 256;
    CUtensorMap ta, tb, tcm;
    if (p.atoms3d) {
        if (int rc = make_fmap_tmap3(&ta, a, B, D, N, dt, bk, BM / BC)) return rc;
    } else {
        if (int rc = make_fmap_tmap(&ta, a, B, D, N, dt, bk)) return rc;
    }
    if (p.b_sw64) {
        if (int rc = make_fmap_tmap3_sw64(&tb, bm, B, D, N, bk, p.BN_mma / 32)) return rc;
    } else if (p.b3d) {
        if (int rc = make_fmap_tmap3(&tb, bm, B, D, N, dt, bk, pair ? p.b_half : p.n_boxes_b)) return rc;
    } else {
        if (int rc = make_fmap_tmap(&tb, bm, B, D, N, dt, bk)) return rc;
    }
    if (p.qbox) {
        if (int rc = make_volume_tmap_atoms(&tcm, level0, B, N, p.BN / 32)) return rc;
    } else if (int rc = make_volume_tmap(&tcm, level0, B, N)) return rc;

    int dev = 0;
    CF_CUDA(cudaGetDevice(&dev));
    static bool opt_in[64] = {};
    if (!opt_in[dev & 63]) {
        CF_CUDA(cudaFuncSetAttribute(corr_tc_kernel<1, false, 1, BK_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        CF_CUDA(cudaFuncSetAttribute(corr_tc_kernel<1, false, 1, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        CF_CUDA(cudaFuncSetAttribute(corr_tc_kernel<2, false, 1, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        CF_CUDA(cudaFuncSetAttribute(corr_tc_kernel<1, true, 1, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        CF_CUDA(cudaFuncSetAttribute(corr_tc_kernel<1, true, 2, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        opt_in[dev & 63] = true;
    }
    const int sms = sm_count();
    if (pair) {
        const int clusters = p.total_tiles < sms / 2 ? p.total_tiles : sms / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * clusters));
        cfg.blockDim = dim3(threads(1));
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CF_CUDA(cudaLaunchKernelEx(&cfg, corr_tc_kernel<2, false, 1, BK>, ta, tb, tcm, p));
        CF_LAUNCH_CHECK("corr_tc_kernel<pair>");
        return CF_OK;
    }
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    if (f16) {
        if (es == 2) corr_tc_kernel<1, true, 2, BK><<<grid, threads(2), smem_bytes, stream>>>(ta, tb, tcm, p);
        else corr_tc_kernel<1, true, 1, BK><<<grid, threads(1), smem_bytes, stream>>>(ta, tb, tcm, p);
        CF_LAUNCH_CHECK("corr_tc_kernel<f16>");
        return CF_OK;
    }
    if (bk == BK_TF32) corr_tc_kernel<1, false, 1, BK_TF32><<<grid, threads(1), smem_bytes, stream>>>(ta, tb, tcm, p);
    else corr_tc_kernel<1, false, 1, BK><<<grid, threads(1), smem_bytes, stream>>>(ta, tb, tcm, p);
    CF_LAUNCH_CHECK("corr_tc_kernel");
    return CF_OK;
}

}  // namespace cf

// debug only (not part of include/cistaflow.h): copies the 32 clock64 stamps of the last launch
extern "C" __attribute__((visibility("default"))) int cf_debug_tc_timeline(unsigned long long *host_out) {
    return cudaMemcpyFromSymbol(host_out, cf::tc::g_tc_timeline, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : -6;
}
