// tc_common.cuh -- shared pieces of the tensor-core (tcgen05 / TMEM / TMA) kernels: PTX wrappers, UMMA descriptors,
// the per-thread k' list, operand-preparation kernels and the exact re-rank kernel.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "refdist.cuh"
#include "select.cuh"

namespace annb {

namespace tc {

constexpr int BM = 128;              // queries per CTA (UMMA M, TMEM lanes)
constexpr int BN = 128;              // database rows per MMA tile (UMMA N, TMEM columns per accumulator)
constexpr int SLAB_BYTES = 128;      // K extent of one smem slab = one 128-byte swizzle atom
constexpr int SLAB_TILE = BM * SLAB_BYTES;  // 16 KiB: 128 rows x 128 B
constexpr int NUM_THREADS = 320;        // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int EPI_THREADS = 256;        // two warps per TMEM lane quarter, each owning one 64-column half of the tile
constexpr int ACC_STAGES = 4;          // accumulator ring in TMEM (4 x 128 columns = all 512)
constexpr int AUX_STAGES = 2;
constexpr uint32_t TMEM_COLS = ACC_STAGES * BN;  // 512

// 3xTF32 (f32 index), bf16 query terms (BF16 index), int8 x int8 -> s32 (SQ8 index), 3xFP16 (f32 index, rows scaled by powers of two)
enum { KIND_TF32X3 = 0, KIND_BF16 = 1, KIND_I8 = 2, KIND_F16X3 = 3 };

struct Params {
    uint64_t nq;
    uint64_t n_rows;
    uint32_t nq_pad;          // rows per query piece in the stacked query operand
    uint32_t n_pad;           // rows per database piece in the stacked database operand
    uint32_t nslab;           // K slabs (KP * elem / 128)
    uint32_t n_stages;        // database ring depth
    uint32_t n_splits;
    uint64_t rows_per_split;  // multiple of BN
    uint32_t a_pieces;        // query terms actually present (bf16 self query: 1)
    uint32_t hybrid;          // bf16, queries in TMEM: the third query term stays in shared memory and is issued as an SS-mode MMA
    uint32_t lo_smem;         // f32, queries in TMEM: only the hi piece; the lo piece stays in shared memory (SS-mode MMA) -- rows of up to 256 elements
    uint32_t stream_q;        // SS mode: query slabs are streamed through the ring with the database slabs (rows too wide for a resident query tile)
    uint32_t wide_k;          // k > k': every list prunes with its OWN threshold only (a value above another list's k'-th may still be in the top-k);
                              // gtau receives the minimum of the lists' final thresholds (the certificate's bound), once per list
    uint32_t strided;         // split y owns the tiles y, y + n_splits, ... (lists interleave over the database) instead of a contiguous range
    const float* aux;         // per database row: L2  v = aux - 2 s  (aux = |x|^2, pad rows +inf);
                              //                   cos v = s * aux    (aux = -1/|x|, pad rows +inf -> 0 * inf = NaN, never selected)
    const float* aux2;        // (KIND_F16X3 operands with per-row scales, cosine: folded into aux; unused by the kernels)
    float db_inv_scale;       // KIND_F16X3, L2: 1 / (the database operand's uniform power-of-two scale)
    const float* q_inv_scale; // KIND_F16X3: per query 1 / (its power-of-two operand scale), [nq_pad]
    uint64_t* part_keys;      // [nq][2 * n_splits][KPRIME] packed (approx value, row); one list per 64-column half
    const void* q_op;         // stacked query operand [a_pieces][nq_pad][kp] (TS mode loads it into TMEM)
    uint32_t kp;              // padded K in elements
    uint32_t* gtau;           // [nq_pad] shared pruning threshold per query (order-preserving image, atomicMin)
    float* dense;             // DENSE mode: the selection values themselves, [nq][dense_ld] (IVF centroid ranking)
    uint32_t dense_ld;
    float* dense_gm;          // DENSE mode, optional: minima of every aligned group of 8 values, [nq][dense_ld / 8] (coarse_select_gm_kernel)
    uint32_t dense_blocked;   // DENSE mode: 1 = the matrix is stored in the blocked layout of dense_piece_index() (coalesced epilogue stores)
    float* dbg;               // optional: CTA (0,0) dumps v of its first tile [BM][BN]
    unsigned long long* dbg_cycles;  // optional: CTA (0,0) wait-cycle counters {total, prod_empty, mma_full, mma_tempty, epi_tfull, epi_slow}
};

// Blocked layout of the DENSE matrix, in float4 pieces (4 consecutive columns of one query row): [query tile of 128][column tile
// of 128][piece 0 .. 31][row in tile 0 .. 127].  An epilogue warp's store instruction (32 rows x one piece) is then 512 contiguous
// bytes instead of 32 half-used sectors 4 * dense_ld bytes apart; readers fetch single pieces (the group-minima select reads two per group).
__host__ __device__ __forceinline__ uint64_t dense_piece_index(uint64_t q, uint32_t i4, uint32_t col_tiles) {
    return (((q >> 7) * col_tiles + (i4 >> 5)) * 32 + (i4 & 31u)) * 128 + (q & 127u);
}

// ----------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
// Role counters (tools/tc_cycles.py, tools/ivf_cycles.py) exist only in the -DANNB_TC_COUNTERS build
// (lib/libannb200_counters.so): the production kernels carry no clock reads on their per-tile chains.
#ifdef ANNB_TC_COUNTERS
constexpr bool TC_COUNTERS = true;
__device__ __forceinline__ long long tc_clock() { return clock64(); }
#else
constexpr bool TC_COUNTERS = false;
__device__ __forceinline__ long long tc_clock() { return 0; }
#endif
__device__ __forceinline__ void mbar_wait_timed(uint64_t* bar, uint32_t parity, long long& acc) {
    const long long t0 = tc_clock();
    mbar_wait(bar, parity);
    acc += tc_clock() - t0;
}
// One lane of a converged warp (elect.sync): lets the compiler keep the surrounding values in uniform registers.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tm, uint64_t* bar, int32_t x, int32_t y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
constexpr uint32_t SMEM_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // 0x40004040
// The 64-bit shared-memory descriptors are handed to the instruction as {low word, high word} pairs assembled inside the
// asm statement.  Their high word is a compile-time constant (SBO, version, swizzle mode); when it reached ptxas as part
// of a 64-bit add with a wide immediate, ptxas 12.9 was seen to drop it on the uniform datapath (the descriptor's high
// word then held an unrelated kernel parameter) -- keeping the halves separate 32-bit values rules that out.
#define ANNB_UMMA_SS(KINDSTR)                                                                                                          \
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t" \
                 "tcgen05.mma.cta_group::1.kind::" KINDSTR " [%0], da, db, %5, p;\n\t}" ::"r"(tmem_c),                                  \
                 "r"(adesc), "r"(SMEM_DESC_HI), "r"(bdesc), "r"(SMEM_DESC_HI), "r"(idesc), "r"(accumulate)                                \
                 : "memory")
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_c, uint32_t adesc, uint32_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == KIND_TF32X3) ANNB_UMMA_SS("tf32");
    else if (KIND == KIND_I8) ANNB_UMMA_SS("i8");
    else ANNB_UMMA_SS("f16");
}
#undef ANNB_UMMA_SS
// 32 lanes x 32 columns of 32-bit accumulators: thread i of the warp receives columns [c, c+32) of TMEM lane base+i.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t r[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 32 columns, completed (load + wait in one statement, see tmem_ld64_sync)
__device__ __forceinline__ void tmem_ld32_sync(uint32_t taddr, uint32_t r[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 64 consecutive columns of this warp's 32 lanes, *completed*: both loads and the tcgen05.wait::ld sit in one asm statement,
// so the compiler cannot move, copy or spill the destination registers while the asynchronous loads are still in flight
// (with separate statements it may, under register pressure, and the late-arriving data then lands in re-used registers).
__device__ __forceinline__ void tmem_ld64_sync(uint32_t taddr, uint32_t r[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%64];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%65];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr), "r"(taddr + 32)
        : "memory");
}
// A operand from TMEM (row m of A = TMEM lane m, K along columns), B from shared memory.
#define ANNB_UMMA_TS(KINDSTR)                                                                                                          \
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"                           \
                 "tcgen05.mma.cta_group::1.kind::" KINDSTR " [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_c),                                \
                 "r"(tmem_a), "r"(bdesc), "r"(SMEM_DESC_HI), "r"(idesc), "r"(accumulate)                                                \
                 : "memory")
template <int KIND>
__device__ __forceinline__ void umma_ts(uint32_t tmem_c, uint32_t tmem_a, uint32_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == KIND_TF32X3) ANNB_UMMA_TS("tf32");
    else if (KIND == KIND_I8) ANNB_UMMA_TS("i8");
    else ANNB_UMMA_TS("f16");
}
#undef ANNB_UMMA_TS
// Store + wait in one asm statement: the store is asynchronous, so its source registers must stay untouched until
// tcgen05.wait::st -- with separate statements the compiler is free to re-use them in between.
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t r[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
        "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n\t"
        "tcgen05.wait::st.sync.aligned;" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
        "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
        "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
// two 32-column stores and one wait in a single statement (see tmem_st32 for why the wait is fused)
__device__ __forceinline__ void tmem_st64(uint32_t taddr, const uint32_t r[64]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
        "%23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33};\n\t"
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%1], {%34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, "
        "%54, %55, %56, %57, %58, %59, %60, %61, %62, %63, %64, %65};\n\t"
        "tcgen05.wait::st.sync.aligned;" ::"r"(taddr), "r"(taddr + 32),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
        "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
        "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(r[32]), "r"(r[33]), "r"(r[34]), "r"(r[35]), "r"(r[36]),
        "r"(r[37]), "r"(r[38]), "r"(r[39]), "r"(r[40]), "r"(r[41]), "r"(r[42]), "r"(r[43]), "r"(r[44]), "r"(r[45]), "r"(r[46]), "r"(r[47]), "r"(r[48]),
        "r"(r[49]), "r"(r[50]), "r"(r[51]), "r"(r[52]), "r"(r[53]), "r"(r[54]), "r"(r[55]), "r"(r[56]), "r"(r[57]), "r"(r[58]), "r"(r[59]), "r"(r[60]),
        "r"(r[61]), "r"(r[62]), "r"(r[63])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major operand tile, 128-byte swizzle, rows 128 B apart, 8-row groups 1024 B apart.
// (field layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor).  Kept as two 32-bit words: the low word carries the start
// address (>> 4, 14 bits -- all descriptor arithmetic in the kernels is on this word) and the LBO field, the high word is
// the constant {SBO = 1024 B, descriptor version 1, SWIZZLE_128B}.
__device__ __forceinline__ uint32_t make_smem_desc(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
// UMMA instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): f32 accumulate, K-major A and B.
__host__ __device__ constexpr uint32_t make_idesc(int kind) {
    const uint32_t fmt = (kind == KIND_TF32X3) ? 2u : (kind == KIND_F16X3 ? 0u : 1u);  // TF32 = 2, F16 = 0, BF16 = 1, signed INT8 = 1 (S8Format)
    const uint32_t cfmt = (kind == KIND_I8) ? 2u : 1u;     // accumulator: F32 = 1, S32 = 2
    return (cfmt << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
}

// ----------------------------------------------------------------------------------------------- per-thread k' list
template <int KP>
struct TopList {
    float v[KP];
    uint32_t i[KP];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j < KP; j++) { v[j] = INFINITY; i[j] = IDX_INVALID; }
    }
    __device__ __forceinline__ float tau() const { return v[KP - 1]; }
    // requires x < tau(); keeps ascending order, earlier entries win ties
    __device__ __forceinline__ void insert(float x, uint32_t idx) {
#pragma unroll
        for (int j = KP - 1; j > 0; j--) {
            const bool shift = v[j - 1] > x;           // element j-1 moves down to j
            const bool here = !shift && (v[j] > x);    // x lands at j
            v[j] = shift ? v[j - 1] : (here ? x : v[j]);
            i[j] = shift ? i[j - 1] : (here ? idx : i[j]);
        }
        if (v[0] > x) { v[0] = x; i[0] = idx; }
    }
};


// Candidate handling of one 64-value half tile for a thread whose tile minimum m beats its threshold tau.
// In the steady state a thread has one such value per tile at most, but several lanes of a warp have one in the same
// tile; so the common case is straight-line code that every such lane executes together: locate the minimum (group via
// the group minima, column via selects -- no dynamically indexed registers, no local memory) and insert it.  Only if a
// second value of the tile also beats the (now tighter) threshold does the thread fall back to the bulk path: stash
// the 64 values in local memory and walk the groups that hold candidates.
template <int KP, int NV = 64>
__device__ __forceinline__ void select_from_tile(TopList<KP>& top, float& tau, const float (&v)[NV], const float (&gm)[NV / 8], float m, uint32_t col0,
                                                 float (&scratch)[NV]) {
    constexpr int NG = NV / 8;
    uint32_t gs = 0;
#pragma unroll
    for (int g = NG - 1; g >= 0; g--) gs = (gm[g] == m) ? static_cast<uint32_t>(g) : gs;   // lowest group holding the minimum
    float s8[8];
#pragma unroll
    for (int j = 0; j < 8; j++) s8[j] = v[j];
#pragma unroll
    for (int g = 1; g < NG; g++) {
        const bool here = gs == static_cast<uint32_t>(g);
#pragma unroll
        for (int j = 0; j < 8; j++) s8[j] = here ? v[g * 8 + j] : s8[j];
    }
    uint32_t js = 0;
#pragma unroll
    for (int j = 7; j >= 0; j--) js = (s8[j] == m) ? static_cast<uint32_t>(j) : js;   // lowest column of the group holding it
    top.insert(m, col0 + gs * 8 + js);
    tau = fminf(tau, top.tau());
    // second-best value of the tile: other groups' minima, and the minimum's own group without it
    float m2 = INFINITY;
#pragma unroll
    for (int g = 0; g < NG; g++) m2 = fminf(m2, gs == static_cast<uint32_t>(g) ? INFINITY : gm[g]);
#pragma unroll
    for (int j = 0; j < 8; j++) m2 = fminf(m2, js == static_cast<uint32_t>(j) ? INFINITY : s8[j]);
    if (m2 < tau) {
        const uint32_t done = gs * 8 + js;
#pragma unroll
        for (int j = 0; j < NV; j++) scratch[j] = v[j];
        uint32_t gmask = 0;
#pragma unroll
        for (int g = 0; g < NG; g++) gmask |= (gm[g] < tau || gs == static_cast<uint32_t>(g)) ? (1u << g) : 0u;
        while (gmask) {
            const int g = __ffs(gmask) - 1;
            gmask &= gmask - 1;
            float t8[8];
#pragma unroll
            for (int j = 0; j < 8; j++) t8[j] = scratch[g * 8 + j];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (t8[j] < tau && static_cast<uint32_t>(g * 8 + j) != done) {
                    top.insert(t8[j], col0 + g * 8 + j);
                    tau = fminf(tau, top.tau());
                }
            }
        }
    }
}


// ----------------------------------------------------------------------------------------------- operand preparation
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// f32 rows (pitch ld_src floats) -> stacked [2][rows_pad][kp] tf32 hi / lo, zero padded.
static __global__ void split_tf32_kernel(const float* __restrict__ src, uint32_t ld_src, uint32_t dim, uint64_t rows, uint64_t rows_pad, uint32_t kp,
                                  float* __restrict__ dst) {
    const uint64_t total = rows_pad * kp;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t r = i / kp;
        const uint32_t c = static_cast<uint32_t>(i - r * kp);
        float hi = 0.f, lo = 0.f;
        if (r < rows && c < dim) {
            const float x = src[r * ld_src + c];
            hi = rna_tf32(x);
            lo = rna_tf32(__fsub_rn(x, hi));
        }
        dst[i] = hi;
        dst[total + i] = lo;
    }
}
// f32 rows -> stacked [2][rows_pad][kp] fp16 hi / lo of the row scaled by a power of two, zero padded; inv_scale[row] = 1 / scale.
// fp16 carries the 11 significant bits of tf32, so hi + lo holds x to 2^-22 |x| exactly like the tf32 split -- but a kind::f16 MMA
// multiplies 16 elements per K step where kind::tf32 multiplies 8: the same three-term product at half the tensor-pipe time.
// What fp16 lacks is exponent range (5 bits); every row is therefore scaled so that its largest element lands in [2^13, 2^14):
// hi never overflows, lo (<= 2^-11 of hi) stays normal for every element within 2^-14 of the row maximum, and smaller elements
// are off by at most 2^-25 against a row maximum of 2^13 -- far below the 2^-22 the split promises relative to the row NORM,
// which is all the certificate's error model uses.  Scales are powers of two: applying and undoing them is exact.
// One warp per row.
// fixed_scale > 0: every row uses this power of two instead of its own (L2 database operands: the certificate's error model is
// relative to the LARGEST row norm anyway, and a uniform scale leaves the epilogue without a per-column scale to undo; elements
// far below the largest one go subnormal -- an absolute error of 2^-25 / scale, i.e. below 2^-38 of the largest element).
static __global__ void split_f16_kernel(const float* __restrict__ src, uint32_t ld_src, uint32_t dim, uint64_t rows, uint64_t rows_pad, uint32_t kp,
                                        __half* __restrict__ dst, float* __restrict__ inv_scale, float sign = 1.0f, float fixed_scale = 0.0f) {   // sign = -1: the negated rows (exact)
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t total = rows_pad * kp;
    for (uint64_t r = blockIdx.x * static_cast<uint64_t>(blockDim.x >> 5) + (threadIdx.x >> 5); r < rows_pad; r += static_cast<uint64_t>(gridDim.x) * (blockDim.x >> 5)) {
        float m = 0.f;
        if (r < rows)
            for (uint32_t c = lane; c < dim; c += 32) m = fmaxf(m, fabsf(src[r * ld_src + c]));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
        int e = 0;
        if (m > 0.f && m < INFINITY) {
            int ex;
            (void)frexpf(m, &ex);                 // m = f * 2^ex, f in [0.5, 1)
            e = min(max(14 - ex, -100), 100);     // m * 2^e in [2^13, 2^14)
        }
        const float scale = fixed_scale > 0.0f ? fixed_scale : ldexpf(1.0f, e);
        for (uint32_t c = lane; c < kp; c += 32) {
            __half h = __float2half_rn(0.f), l = h;
            if (r < rows && c < dim) {
                const float xs = __fmul_rn(src[r * ld_src + c], scale * sign);
                h = __float2half_rn(xs);
                l = __float2half_rn(__fsub_rn(xs, __half2float(h)));
            }
            dst[r * kp + c] = h;
            dst[total + r * kp + c] = l;
        }
        if (lane == 0 && inv_scale != nullptr) inv_scale[r] = fixed_scale > 0.0f ? 1.0f / fixed_scale : ldexpf(1.0f, -e);
    }
}
// largest finite |value| of an f32 array (bit pattern; non-negative floats order like their bits)
static __global__ void absmax_kernel(const float* __restrict__ p, uint64_t n, uint32_t* __restrict__ out_bits) {
    uint32_t m = 0;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const float v = fabsf(p[i]);
        if (v < INFINITY) m = max(m, __float_as_uint(v));
    }
    atomicMax(out_bits, m);
}
// Database operand of the flat f32 cosine kernel (flat_tc_kernel, UNIT): every row divided by its index norm (one f32 division per
// element: 2^-24 relative, budgeted in tc_cert_eps), times the uniform scale 2^13, as stacked fp16 hi / lo pieces.  Unit rows have
// their largest element in [2^13 / sqrt(dim), 2^13]: hi never overflows and the pieces hold the row to far better than 2^-22 of
// its norm, as the per-row scales of split_f16_kernel do for un-normalised rows.  Rows past `rows` (tile padding) and rows whose
// norm is zero or not finite carry a NaN in the first element of the hi piece: their whole accumulator column is NaN, which the
// epilogue's min tree and threshold test ignore -- the role of the NaN row constant in the other operand forms.  One warp per row.
static __global__ void split_f16_unit_kernel(const float* __restrict__ src, uint32_t ld_src, uint32_t dim, uint64_t rows, uint64_t rows_pad, uint32_t kp,
                                             const float* __restrict__ norms, __half* __restrict__ dst) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t total = rows_pad * kp;
    for (uint64_t r = blockIdx.x * static_cast<uint64_t>(blockDim.x >> 5) + (threadIdx.x >> 5); r < rows_pad; r += static_cast<uint64_t>(gridDim.x) * (blockDim.x >> 5)) {
        const float nrm = r < rows ? norms[r] : 0.f;
        const bool ok = nrm > 0.f && nrm < INFINITY;
        for (uint32_t c = lane; c < kp; c += 32) {
            __half h = __float2half_rn(0.f), l = h;
            if (ok && c < dim) {
                const float xs = __fmul_rn(__fdiv_rn(src[r * ld_src + c], nrm), 8192.0f);
                h = __float2half_rn(xs);
                l = __float2half_rn(__fsub_rn(xs, __half2float(h)));
            }
            if (!ok && c == 0) h = __ushort_as_half(static_cast<unsigned short>(0x7E00));   // quiet NaN
            dst[r * kp + c] = h;
            dst[total + r * kp + c] = l;
        }
    }
}
// inv_scale[r] = 1 / (power-of-two scale that brings the largest element of row r into [2^13, 2^14)) -- the scale split_f16_kernel
// would choose; the IVF scan converts its f32 slabs inside the kernel and only needs the scales.  One warp per row.
static __global__ void row_inv_scale_kernel(const float* __restrict__ src, uint32_t ld_src, uint32_t dim, uint64_t rows, uint64_t rows_pad, float* __restrict__ inv_scale) {
    const uint32_t lane = threadIdx.x & 31u;
    for (uint64_t r = blockIdx.x * static_cast<uint64_t>(blockDim.x >> 5) + (threadIdx.x >> 5); r < rows_pad; r += static_cast<uint64_t>(gridDim.x) * (blockDim.x >> 5)) {
        float m = 0.f;
        if (r < rows)
            for (uint32_t c = lane; c < dim; c += 32) m = fmaxf(m, fabsf(src[r * ld_src + c]));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
        int e = 0;
        if (m > 0.f && m < INFINITY) {
            int ex;
            (void)frexpf(m, &ex);
            e = min(max(14 - ex, -100), 100);
        }
        if (lane == 0) inv_scale[r] = ldexpf(1.0f, -e);
    }
}
static __global__ void fill_f32_value_kernel(float* __restrict__ p, uint64_t n, float v) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i < n) p[i] = v;
}
// aux[i] *= s[i] (cosine row constants of the 3xFP16 kernel carry the row's inverse operand scale)
static __global__ void mul_rows_kernel(float* __restrict__ aux, const float* __restrict__ s, uint64_t n) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i < n) aux[i] *= s[i];
}
// f32 queries -> stacked [3][rows_pad][kp] bf16 terms q0 + q1 + q2 (each RNE), zero padded.
static __global__ void split_bf16x3_kernel(const float* __restrict__ src, uint32_t ld_src, uint32_t dim, uint64_t rows, uint64_t rows_pad, uint32_t kp,
                                    __nv_bfloat16* __restrict__ dst) {
    const uint64_t total = rows_pad * kp;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t r = i / kp;
        const uint32_t c = static_cast<uint32_t>(i - r * kp);
        __nv_bfloat16 b0 = __float2bfloat16_rn(0.f), b1 = b0, b2 = b0;
        if (r < rows && c < dim) {
            const float x = src[r * ld_src + c];
            b0 = __float2bfloat16_rn(x);
            const float r1 = __fsub_rn(x, __bfloat162float(b0));
            b1 = __float2bfloat16_rn(r1);
            b2 = __float2bfloat16_rn(__fsub_rn(r1, __bfloat162float(b1)));
        }
        dst[i] = b0;
        dst[total + i] = b1;
        dst[2 * total + i] = b2;
    }
}
// bf16 rows (pitch ld_src elements) -> [rows_pad][kp] bf16, zero padded (database operand / bf16 self queries).
static __global__ void pad_bf16_kernel(const uint16_t* __restrict__ src, uint32_t ld_src, uint32_t dim, uint64_t rows, uint64_t rows_pad, uint32_t kp,
                                uint16_t* __restrict__ dst) {
    const uint64_t total = rows_pad * kp;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t r = i / kp;
        const uint32_t c = static_cast<uint32_t>(i - r * kp);
        dst[i] = (r < rows && c < dim) ? src[r * ld_src + c] : static_cast<uint16_t>(0);
    }
}
// Epilogue constants per database row.  L2: (-2, |x|^2) with |x|^2 of the stored (possibly bf16-rounded) row;
// cosine: (-1/norm, 0) with the index norm (f32 norm of the un-rounded row, as the reference divides by it).
// max over rows of aux (L2: |x|^2, all >= 0): non-negative floats order like their bit patterns
static __global__ void aux_max_kernel(const float* __restrict__ aux, uint64_t n, uint32_t* __restrict__ out_bits) {
    uint32_t m = 0;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const float v = aux[i];
        if (v == v && v < INFINITY && v > 0.f) m = max(m, __float_as_uint(v));
    }
    atomicMax(out_bits, m);
}
static __global__ void aux_kernel(const uint8_t* __restrict__ rows, uint32_t row_bytes, int rt, uint32_t dim, const float* __restrict__ norms,
                                  const int32_t* __restrict__ norms_i, int cosine, uint64_t n, uint64_t n_pad_total, float* __restrict__ aux) {
    // rt: 0 f32 rows, 1 bf16 rows, 2 int8 codes.  L2: aux = |x|^2 (of the stored, possibly rounded / quantised row);
    // cosine: aux = -1/norm with the index norm (f32 norm of the un-rounded row; sqrt of the code norm for SQ8, 0 if that is 0).
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n_pad_total) return;
    float o = INFINITY;
    if (i < n) {
        if (cosine) {
            if (rt == 2) {
                const float nn = sqrtf(static_cast<float>(norms_i[i]));
                o = nn > 0.f ? -1.0f / nn : 0.f;
            } else {
                o = -1.0f / norms[i];
            }
        } else {
            const uint8_t* r = rows + i * row_bytes;
            if (rt == 2) {
                int32_t s = 0;
                for (uint32_t e = 0; e < dim; e++) {
                    const int32_t x = reinterpret_cast<const int8_t*>(r)[e];
                    s += x * x;
                }
                o = static_cast<float>(s);
            } else {
                double s = 0.0;   // f64, rounded once: the certificates budget 2^-24 |x|^2 for this constant
                for (uint32_t e = 0; e < dim; e++) {
                    const double x = rt == 1 ? bf16_bits_to_f32(reinterpret_cast<const uint16_t*>(r)[e]) : reinterpret_cast<const float*>(r)[e];
                    s += x * x;
                }
                o = static_cast<float>(s);
            }
        }
    }
    aux[i] = o;
}
// rows of src_bytes -> rows of dst_bytes (>= src_bytes), zero padded
static __global__ void repitch_rows_kernel(const uint8_t* __restrict__ src, uint32_t src_bytes, uint8_t* __restrict__ dst, uint32_t dst_bytes, uint64_t n) {
    const uint64_t total = n * dst_bytes;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t r = i / dst_bytes;
        const uint32_t c = static_cast<uint32_t>(i - r * dst_bytes);
        dst[i] = c < src_bytes ? src[r * src_bytes + c] : 0;
    }
}
// int8 codes (pitch ld_src bytes) -> [rows_pad][kp] zero padded (database operand / encoded queries of the SQ8 index).
static __global__ void pad_i8_kernel(const int8_t* __restrict__ src, uint32_t ld_src, uint32_t dim, uint64_t rows, uint64_t rows_pad, uint32_t kp,
                                     int8_t* __restrict__ dst) {
    const uint64_t total = rows_pad * kp;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t r = i / kp;
        const uint32_t c = static_cast<uint32_t>(i - r * kp);
        dst[i] = (r < rows && c < dim) ? src[r * ld_src + c] : static_cast<int8_t>(0);
    }
}

// ----------------------------------------------------------------------------------------------- exact re-rank + merge
struct RerankParams {
    const uint64_t* part_keys;  // [nq][parts][kp] approximate keys
    uint32_t parts, kp, k_eff, k_out, nsort;
    uint32_t n_exact;           // wide-k mode (k > k'): candidates recomputed exactly per query (power of two > 64, exact keys behind keys[nsort]); 0 otherwise
    uint64_t nq;
    const uint8_t* rows;  // index rows in the index dtype
    uint32_t row_bytes;
    const float* row_norms;
    const int32_t* row_norms_i;  // SQ8 cosine
    const uint8_t* queries;  // prepared queries (f32 padded rows, or bf16 rows for self queries)
    uint32_t q_bytes;
    uint32_t dim;
    int bf16_self;
    uint64_t id_base;
    const uint32_t* parts_used;  // optional [nq]: only the first parts_used[q] * part_mult lists of a query hold data (IVF probe ranks)
    uint32_t part_mult;
    const uint64_t* id_map;      // optional: id = id_map[row] (IVF original ids) instead of row + id_base
    const uint64_t* row_map;     // optional: output row of query q
    // Coverage certificate: every row that was NOT re-ranked has an approximate value >= the k'-th merged approximate
    // value A.  If A (mapped back to distance units) exceeds the k-th exact distance by more than the error bound of the
    // approximate values, no such row can belong to the exact top-k: the result is provably the reference's.
    float cert_eps;              // relative error bound of the pre-selection values (0 = certificate off)
    float xnorm_max;             // L2: largest stored-row norm
    uint32_t* uncert_count;      // number of queries that could not be certified
    uint32_t* uncert_list;       // their indices (capacity nq)
    const uint32_t* gtau;        // optional [nq]: final shared pruning threshold of the scan (ordered image); bounds the k'-th merged value
    float* out_bound;            // optional [nq] (shard mode): distance bound of the rows that were not re-ranked, instead of a local verdict
    uint64_t* out_ids;
    float* out_dist;
    uint32_t* out_counts;
};

// One CTA per query: merge the per-split candidate lists by approximate value, keep the best kp, recompute their
// distances exactly (reference order, bit-identical to the CPU path), order by (distance, id) and emit k.
template <int RT, int QT, int MET>
__global__ void __launch_bounds__(128) rerank_kernel(RerankParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem);
    __shared__ uint64_t exact64[64];
    const bool wide = p.n_exact > 64u;
    uint64_t* exact = wide ? keys + p.nsort : exact64;
    const uint64_t q = blockIdx.x;
    const uint32_t parts_q = p.parts_used ? min(p.parts, p.parts_used[q] * p.part_mult) : p.parts;
    const uint32_t total = parts_q * p.kp;
    const uint32_t nsort = min(p.nsort, next_pow2(max(total, 64u)));
    const uint64_t* src = p.part_keys + q * (static_cast<uint64_t>(p.parts) * p.kp);
    uint32_t n_cand = 0;   // candidates in keys[] after compaction (gtau path)
    if (p.gtau != nullptr) {
        // Only keys at or below the query's shared pruning threshold can be among the k' best of the merged lists (the
        // threshold is some single list's k'-th value): compact those -- typically a few dozen of the parts * k' slots --
        // and sort just them.
        __shared__ uint32_t s_n;
        const uint32_t thr = p.gtau[q];
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        for (uint32_t i0 = threadIdx.x; i0 < total; i0 += 4 * blockDim.x) {   // four loads in flight per thread: one round trip per 512 slots
            uint64_t kk[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { const uint32_t i = i0 + j * blockDim.x; kk[j] = i < total ? __ldg(src + i) : KEY_SENTINEL; }
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (kk[j] != KEY_SENTINEL && static_cast<uint32_t>(kk[j] >> 32) <= thr) keys[atomicAdd(&s_n, 1u)] = kk[j];
        }
        __syncthreads();
        const uint32_t n = s_n;
        n_cand = n;
        const uint32_t nsort2 = min(nsort, next_pow2(max(n, 64u)));
        if (nsort2 <= 128u) {
            // a few dozen candidates (IVF): one warp sorts them in registers -- no block barrier per bitonic stage.  (Flat searches
            // leave ~200 keys below the smallest list threshold; a single warp sorting 256 keys in registers was measured slower than
            // the block-wide sort: 0.201 against 0.186 ms per 10k queries.)
            if (threadIdx.x < 32) {
                uint64_t kreg[4];
#pragma unroll
                for (int j = 0; j < 4; j++) { const uint32_t i = threadIdx.x + 32u * j; kreg[j] = i < n ? keys[i] : KEY_SENTINEL; }
                warp_bitonic_sort_regs<4>(kreg, threadIdx.x);
#pragma unroll
                for (int j = 0; j < 4; j++) { const uint32_t i = threadIdx.x + 32u * j; if (i < nsort2) keys[i] = kreg[j]; }
            }
            __syncthreads();
        } else {
        for (uint32_t i = n + threadIdx.x; i < nsort2; i += blockDim.x) keys[i] = KEY_SENTINEL;
        __syncthreads();
        bitonic_sort_keys<true>(keys, nsort2, threadIdx.x, blockDim.x);
        }
    } else {
        for (uint32_t i = threadIdx.x; i < nsort; i += blockDim.x) keys[i] = (i < total) ? src[i] : KEY_SENTINEL;
        __syncthreads();
        bitonic_sort_keys<true>(keys, nsort, threadIdx.x, blockDim.x);
    }
    const uint64_t orow = p.row_map ? p.row_map[q] : q;
    const uint8_t* qv = p.queries + q * p.q_bytes;
    // Exact distances of candidates keys[0 .. count) in the reference's arithmetic, written to exact[0 .. 64) (sentinels
    // beyond count).  Eight threads share one candidate: thread l of a group owns SIMD lane l of the reference's 8-lane
    // accumulation (src/utils/dist.rs:306-330, 3615-3636) -- element 8c + l of every 8-element chunk, in chunk order, with
    // the same rounding steps -- and the group folds its eight partial sums with the reference's own horizontal-add tree
    // (shuffles), then thread 0 appends the scalar tail and finishes.  Same bits as one thread walking the row
    // (refdist.cuh), an eighth of the dependent chain.  16 candidates per round of the 128 threads.
    __shared__ float s_qn;
    __shared__ int32_t s_qs;
    if (threadIdx.x == 96) {   // query scalars, once per CTA (a sequential fold in the reference: one thread, off the other warps' path)
        float qn = 1.0f;
        int32_t qs = 0;
        if constexpr (RT == 2) {
            for (uint32_t e = 0; e < p.dim; e++) {
                const int32_t v = reinterpret_cast<const int8_t*>(qv)[e];
                qs += v * v;
            }
        } else if (MET == MET_COS) {
            qn = seq_norm<(QT == QT_F32) ? 4 : 2>(qv, p.dim);
            if (p.bf16_self) qn = round_to_bf16(qn);
        }
        s_qn = qn;
        s_qs = qs;
    }
    // |q|^2 in f64 for the certificate's value -> distance map (any summation order: the rounding is 2^-53-sized)
    __shared__ double s_qn2w[4];
    {
        double part = 0.0;
        for (uint32_t e = threadIdx.x; e < p.dim; e += blockDim.x) {
            double x;
            if constexpr (QT == QT_I8) x = static_cast<double>(reinterpret_cast<const int8_t*>(qv)[e]);
            else x = static_cast<double>(load1<(QT == QT_F32) ? 4 : 2>(qv, e));
            part += x * x;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, off);
        if ((threadIdx.x & 31u) == 0) s_qn2w[threadIdx.x >> 5] = part;
    }
    __syncthreads();
    auto exact_range = [&](uint32_t count, uint32_t base = 0u) {   // candidates keys[base .. base + count), count <= 64 -> exact[base .. base + 64)
        // up to four candidates per 8-thread group (j = grp, grp + 16, grp + 32, grp + 48) with independent accumulators: the row
        // gathers of all of them are in flight together -- one DRAM round trip for a 64-candidate second chance instead of four
        const uint32_t l = threadIdx.x & 7u, grp = threadIdx.x >> 3;
        const uint32_t rounds = min(4u, (count + 15u) >> 4);     // uniform over the CTA
        uint32_t idx[4];
        const uint8_t* row[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint32_t j = r * 16 + grp;
            const uint64_t key = (static_cast<uint32_t>(r) < rounds && j < count) ? keys[base + j] : KEY_SENTINEL;
            idx[r] = key_idx(key);
            row[r] = p.rows + static_cast<uint64_t>(idx[r] != IDX_INVALID ? idx[r] : 0u) * p.row_bytes;
        }
        uint64_t out[4] = {KEY_SENTINEL, KEY_SENTINEL, KEY_SENTINEL, KEY_SENTINEL};
        if constexpr (RT == 2) {
            // SQ8: exact code-space integers (src/utils/dist.rs:5015-5077); any summation order gives the same value
            int32_t dot[4] = {0, 0, 0, 0}, xx[4] = {0, 0, 0, 0};
            for (uint32_t c = l; c < (p.dim + 15u) / 16u; c += 8) {
                const int4 y = *reinterpret_cast<const int4*>(qv + c * 16);
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    if (static_cast<uint32_t>(r) < rounds) {
                        const int4 x = *reinterpret_cast<const int4*>(row[r] + c * 16);
                        xx[r] = __dp4a(x.x, x.x, xx[r]); xx[r] = __dp4a(x.y, x.y, xx[r]); xx[r] = __dp4a(x.z, x.z, xx[r]); xx[r] = __dp4a(x.w, x.w, xx[r]);
                        dot[r] = __dp4a(x.x, y.x, dot[r]); dot[r] = __dp4a(x.y, y.y, dot[r]); dot[r] = __dp4a(x.z, y.z, dot[r]); dot[r] = __dp4a(x.w, y.w, dot[r]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 4; r++) {
                if (static_cast<uint32_t>(r) < rounds) {
#pragma unroll
                    for (int off = 4; off > 0; off >>= 1) {
                        dot[r] += __shfl_down_sync(0xFFFFFFFFu, dot[r], off, 8);
                        xx[r] += __shfl_down_sync(0xFFFFFFFFu, xx[r], off, 8);
                    }
                    if (idx[r] != IDX_INVALID) out[r] = make_key(finish_i8<MET>(dot[r], xx[r], s_qs, (MET == MET_COS) ? p.row_norms_i[idx[r]] : 0), idx[r]);
                }
            }
        } else {
            constexpr int RELEM = (RT == 0) ? 4 : 2, QELEM = (QT == QT_F32) ? 4 : 2;
            constexpr bool FMA = (RELEM == 2);
            const uint32_t chunks = p.dim >> 3;
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 2
            for (uint32_t c = 0; c < chunks; c++) {
                const float y = load1<QELEM>(qv, c * 8 + l);
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    if (static_cast<uint32_t>(r) < rounds) {
                        const float x = load1<RELEM>(row[r], c * 8 + l);
                        if (MET == MET_L2) {
                            const float d = __fsub_rn(x, y);
                            acc[r] = FMA ? __fmaf_rn(d, d, acc[r]) : __fadd_rn(acc[r], __fmul_rn(d, d));
                        } else {
                            acc[r] = FMA ? __fmaf_rn(x, y, acc[r]) : __fadd_rn(acc[r], __fmul_rn(x, y));
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 4; r++) {
                if (static_cast<uint32_t>(r) < rounds) {
                    // s_l = a_l + a_(l+4);  wide: (s0 + s2) + (s1 + s3);  hsum_f32_avx2: (s0 + s1) + (s2 + s3)
                    const float s4 = __fadd_rn(acc[r], __shfl_down_sync(0xFFFFFFFFu, acc[r], 4, 8));
                    float sum;
                    if (FMA) {
                        const float u = __fadd_rn(s4, __shfl_down_sync(0xFFFFFFFFu, s4, 1, 8));
                        sum = __fadd_rn(u, __shfl_down_sync(0xFFFFFFFFu, u, 2, 8));
                    } else {
                        const float t = __fadd_rn(s4, __shfl_down_sync(0xFFFFFFFFu, s4, 2, 8));
                        sum = __fadd_rn(t, __shfl_down_sync(0xFFFFFFFFu, t, 1, 8));
                    }
                    if (l == 0 && idx[r] != IDX_INVALID) {
                        for (uint32_t e = chunks * 8; e < p.dim; e++) {   // scalar tail: `sum += d * d` (not fused)
                            const float x = load1<RELEM>(row[r], e), y = load1<QELEM>(qv, e);
                            if (MET == MET_L2) {
                                const float d = __fsub_rn(x, y);
                                sum = __fadd_rn(sum, __fmul_rn(d, d));
                            } else {
                                sum = __fadd_rn(sum, __fmul_rn(x, y));
                            }
                        }
                        out[r] = make_key(finish_fp<MET>(sum, s_qn, (MET == MET_COS) ? p.row_norms[idx[r]] : 1.0f), idx[r]);
                    }
                }
            }
        }
        if (l == 0) {
#pragma unroll
            for (int r = 0; r < 4; r++) exact[base + r * 16 + grp] = out[r];
        }
    };
    // coverage test: every row that was not re-ranked has an approximate value >= a_thr; is the k-th exact distance safely
    // below the distance that value stands for?  (one thread, f64: the map itself must not add rounding of its own)
    // bound_of(a): every row whose approximate value is >= a has a reference-order distance ABOVE this bound
    auto bound_of = [&](float a_thr) -> double {
        const double qn2 = (s_qn2w[0] + s_qn2w[1]) + (s_qn2w[2] + s_qn2w[3]);
        if (MET == MET_L2) {
            // approx value = |x|^2 - 2 q.x = dist - |q|^2 ; error <= eps * (|q| + |x|max)^2
            const double s = sqrt(qn2) + static_cast<double>(p.xnorm_max);
            return static_cast<double>(a_thr) + qn2 - static_cast<double>(p.cert_eps) * s * s;
        }
        // approx value = -q.x / |x| = (dist - 1) * |q| with the reference's own |q| (sequential fold, bf16-rounded for
        // bf16 self queries; sqrt of the integer norm for SQ8) ; error <= eps
        double qn = sqrt(qn2);
        if constexpr (QT != QT_I8) qn = static_cast<double>(s_qn);
        return qn > 0.0 ? (static_cast<double>(a_thr) / qn + 1.0 - static_cast<double>(p.cert_eps)) : static_cast<double>(INFINITY);
    };
    auto covered = [&](float a_thr, float dk) -> bool { return bound_of(a_thr) > static_cast<double>(dk); };
    __shared__ int s_extend;
    if (wide) {
        // k > k' (wide-k mode, gtau path only).  No list shared its threshold, so G = gtau[q] is the minimum of the lists' final
        // thresholds: a row that is in no list has a value >= G, and every scanned row with a value below G sits in the compacted
        // set.  All of them (up to n_exact, in approximate order) are recomputed exactly; the query is certified iff its k-th exact
        // distance lies below the distance bound of G (or of the first candidate that was not recomputed).
        const uint32_t ne = min(n_cand, p.n_exact);
        for (uint32_t base = 0; base < p.n_exact; base += 64) {
            if (base < ne) exact_range(min(64u, ne - base), base);
            else if (threadIdx.x < 64) exact[base + threadIdx.x] = KEY_SENTINEL;
        }
        __syncthreads();
        bitonic_sort_keys<true>(exact, p.n_exact, threadIdx.x, blockDim.x);
        if (threadIdx.x == 0) {
            const uint32_t g = p.gtau[q];
            double b = static_cast<double>(INFINITY);                       // nothing was ever pruned: every row is a candidate
            if (n_cand > ne) b = bound_of(key_dist(keys[ne]));
            else if (g != 0xFFFFFFFFu) b = bound_of(ordered_to_f32(g));
            const uint64_t d_key = exact[p.k_eff - 1];
            const bool ok = key_idx(d_key) != IDX_INVALID ? (b > static_cast<double>(key_dist(d_key))) : (g == 0xFFFFFFFFu && n_cand <= ne);
            if (p.out_bound != nullptr) p.out_bound[q] = (p.cert_eps > 0.f) ? __double2float_rd(b) : INFINITY;
            else if (!ok && p.uncert_count != nullptr && p.cert_eps > 0.f) p.uncert_list[atomicAdd(p.uncert_count, 1u)] = static_cast<uint32_t>(q);
        }
        __syncthreads();
    } else {
    const bool certify = (p.uncert_count != nullptr || p.out_bound != nullptr) && p.cert_eps > 0.f;
    // Outcome of the first test, predicted from the approximate values alone (no memory traffic): the k'-th merged value must
    // clear the k-th candidate's value by the certificate's margin.  A query predicted to fail (on squared distances most do:
    // |x|^2 dwarfs the neighbour gaps) skips the k'-candidate pass and goes straight to the 64-candidate one, whose certificate
    // is the stronger of the two anyway -- one row-gather round trip and one sort per query instead of two.  A wrong prediction
    // costs time only: both passes certify on their own terms.
    bool direct = false;
    if (certify && p.gtau != nullptr && n_cand > p.kp && p.k_eff <= p.kp) {
        const float a_kp = key_dist(keys[p.kp - 1]), a_k = key_dist(keys[p.k_eff - 1]);
        const double qn2 = (s_qn2w[0] + s_qn2w[1]) + (s_qn2w[2] + s_qn2w[3]);
        double d_pred;       // distance the k-th candidate's approximate value stands for
        if (MET == MET_L2) d_pred = static_cast<double>(a_k) + qn2;
        else { double qn = sqrt(qn2); if constexpr (QT != QT_I8) qn = static_cast<double>(s_qn); d_pred = qn > 0.0 ? static_cast<double>(a_k) / qn + 1.0 : 0.0; }
        direct = !(bound_of(a_kp) > d_pred);
    }
    if (threadIdx.x == 0) s_extend = direct ? 1 : 0;
    if (threadIdx.x == 0 && p.out_bound != nullptr) p.out_bound[q] = INFINITY;
    if (!direct) {
    exact_range(p.kp);
    __syncthreads();
    if (threadIdx.x < 32) { uint64_t er[2] = {exact[threadIdx.x], exact[threadIdx.x + 32]}; warp_bitonic_sort_regs<2>(er, threadIdx.x); exact[threadIdx.x] = er[0]; exact[threadIdx.x + 32] = er[1]; }
    __syncthreads();
    // Shard mode (out_bound != nullptr): the certificate is not decided here.  The kernel reports the bound below which
    // this shard's un-re-ranked rows cannot lie; the caller tests it against the k-th distance of the MERGED result
    // (annb_shard_check_dev) -- a shard holding only far lists of a query need not be exact about candidates that cannot
    // reach the global top-k.
    if (threadIdx.x == 0 && certify) {
        const uint64_t a_key = keys[p.kp - 1];                       // k'-th merged approximate key (sentinel: every row was re-ranked)
        const uint64_t d_key = exact[p.k_eff - 1];                   // k-th exact key (sentinel: fewer than k rows exist)
        if (key_idx(a_key) != IDX_INVALID && p.out_bound != nullptr) p.out_bound[q] = __double2float_rd(bound_of(key_dist(a_key)));
        if (key_idx(a_key) != IDX_INVALID && key_idx(d_key) != IDX_INVALID && !covered(key_dist(a_key), key_dist(d_key))) {
            // Second chance before the exact fallback: the compacted candidate set holds every scanned row whose value is
            // at or below the query's final pruning threshold G (every rejected row is >= G, which is usually well above
            // the k'-th merged value).  Re-rank up to 64 of them and test against G (or the 65th value).
            if (p.gtau != nullptr && n_cand > p.kp) s_extend = 1;
            else if (p.out_bound == nullptr) p.uncert_list[atomicAdd(p.uncert_count, 1u)] = static_cast<uint32_t>(q);
        }
    }
    }   // !direct
    __syncthreads();
    if (s_extend) {
        exact_range(min(n_cand, 64u));
        __syncthreads();
        if (threadIdx.x < 32) { uint64_t er[2] = {exact[threadIdx.x], exact[threadIdx.x + 32]}; warp_bitonic_sort_regs<2>(er, threadIdx.x); exact[threadIdx.x] = er[0]; exact[threadIdx.x + 32] = er[1]; }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint64_t d_key = exact[p.k_eff - 1];
            double b = static_cast<double>(INFINITY);
            if (n_cand > 64) b = bound_of(key_dist(keys[64]));
            else if (p.gtau[q] != 0xFFFFFFFFu) b = bound_of(ordered_to_f32(p.gtau[q]));   // 0xFFFFFFFF: nothing was ever pruned
            if (p.out_bound != nullptr) p.out_bound[q] = __double2float_rd(b);
            else if (!(b > static_cast<double>(key_dist(d_key)))) p.uncert_list[atomicAdd(p.uncert_count, 1u)] = static_cast<uint32_t>(q);
        }
    }
    }   // !wide
    const uint32_t exact_cap = wide ? p.n_exact : 64u;
    uint32_t valid = 0;
    for (uint32_t j = threadIdx.x; j < p.k_out; j += blockDim.x) {
        uint64_t key = (j < p.k_eff && j < exact_cap) ? exact[j] : KEY_SENTINEL;
        uint64_t id = 0xFFFFFFFFFFFFFFFFull;
        float d = INFINITY;
        if (key_idx(key) != IDX_INVALID) {
            id = p.id_map ? p.id_map[key_idx(key)] : static_cast<uint64_t>(key_idx(key)) + p.id_base;
            d = key_dist(key);
            valid++;
        }
        p.out_ids[orow * p.k_out + j] = id;
        if (p.out_dist) p.out_dist[orow * p.k_out + j] = d;
    }
    if (p.out_counts) {
        __shared__ uint32_t s_cnt;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        if (valid) atomicAdd(&s_cnt, valid);
        __syncthreads();
        if (threadIdx.x == 0) p.out_counts[orow] = s_cnt;
    }
}

}  // namespace tc

}  // namespace annb
