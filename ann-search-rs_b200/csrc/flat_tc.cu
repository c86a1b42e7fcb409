// flat_tc.cu -- tensor-core flat search: tcgen05.mma into TMEM, operands fed by TMA, top-k' selection fused
// into the TMEM epilogue (the [nq x n] distance matrix never exists), then an exact re-rank of the k'
// survivors in the reference's own floating-point order (refdist.cuh).
//
// Replaces, for the flat f32 / bf16 indices, the reference's distance-matrix kernels + separate top-k pass
// (src/gpu/dist_gpu.rs:79-488 euclidean/cosine_tiled{,_reg}; :553-613 extract_topk; src/gpu/topk_gpu.rs:992-1237).
//
// Kernel shape (one CTA = 128 queries x one database split, 192 threads, 1 CTA / SM):
//   warp 0      TMA producer: query tile once, then a ring of database K-slabs (128 rows x 128 B, SWIZZLE_128B)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer, accumulators double buffered in TMEM
//   warps 2..5  epilogue: tcgen05.ld 32 columns at a time, v = fma(s, a[col], b[col]) (norm terms fused),
//               min-tree + threshold test; rare insert into a register-resident sorted k' list per query
//   f32 index : 3xTF32 split precision  s = Qhi.Xhi + Qlo.Xhi + Qhi.Xlo   (kind::tf32, hi/lo rounded with cvt.rna)
//   bf16 index: f32 queries split into three bf16 terms, s = (q0 + q1 + q2).X  (kind::f16), X = the stored bf16 rows
#include <cuda.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "flat_tc.hpp"
#include "index.hpp"
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

enum { KIND_TF32X3 = 0, KIND_BF16 = 1 };

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
    const float* aux;         // per database row: L2  v = aux - 2 s  (aux = |x|^2, pad rows +inf);
                              //                   cos v = s * aux    (aux = -1/|x|, pad rows +inf -> 0 * inf = NaN, never selected)
    uint64_t* part_keys;      // [nq][2 * n_splits][KPRIME] packed (approx value, row); one list per 64-column half
    const void* q_op;         // stacked query operand [a_pieces][nq_pad][kp] (TS mode loads it into TMEM)
    uint32_t kp;              // padded K in elements
    uint32_t* gtau;           // [nq_pad] shared pruning threshold per query (order-preserving image, atomicMin)
    float* dbg;               // optional: CTA (0,0) dumps v of its first tile [BM][BN]
    unsigned long long* dbg_cycles;  // optional: CTA (0,0) wait-cycle counters {total, prod_empty, mma_full, mma_tempty, epi_tfull, epi_slow}
};

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
__device__ __forceinline__ void mbar_wait_timed(uint64_t* bar, uint32_t parity, long long& acc) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
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
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == KIND_TF32X3)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
                     "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
                     "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                     : "memory");
}
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
// A operand from TMEM (row m of A = TMEM lane m, K along columns), B from shared memory.
template <int KIND>
__device__ __forceinline__ void umma_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == KIND_TF32X3)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_c),
                     "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_c),
                     "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
                     : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t r[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
        "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
        "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
        "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major operand tile, 128-byte swizzle, rows 128 B apart, 8-row groups 1024 B apart.
// (field layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    return static_cast<uint64_t>((saddr >> 4) & 0x3FFFu) | (1ull << 16)  // LBO (unused for swizzled K-major)
           | (static_cast<uint64_t>(1024 >> 4) << 32)                    // SBO = 1024 B
           | (1ull << 46)                                                // descriptor version (Blackwell)
           | (2ull << 61);                                               // SWIZZLE_128B
}
// UMMA instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): f32 accumulate, K-major A and B.
__host__ __device__ constexpr uint32_t make_idesc(int kind) {
    const uint32_t fmt = (kind == KIND_TF32X3) ? 2u : 1u;  // TF32 = 2, BF16 = 1
    return (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
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

// ----------------------------------------------------------------------------------------------- the kernel
// TS = the query operand lives in TMEM (columns [0, 256): hi then lo) instead of shared memory: the MMAs read only the
// database slab from shared memory (half the operand bandwidth) and the whole 227 KB becomes database ring.
template <int KIND, int KP, int MET, bool TS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
flat_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_x, const Params p) {
    constexpr int NA = (KIND == KIND_TF32X3) ? 2 : 3;  // stacked query pieces
    constexpr int NB = (KIND == KIND_TF32X3) ? 2 : 1;  // stacked database pieces
    constexpr int ELEM = (KIND == KIND_TF32X3) ? 4 : 2;
    constexpr int SLAB_ELEMS = SLAB_BYTES / ELEM;      // 32 tf32 / 64 bf16 per slab row
    constexpr int KSTEPS = 4;                          // 128 B / 32 B per UMMA K step (8 tf32 / 16 bf16)
    constexpr int NACC = TS ? 2 : ACC_STAGES;          // accumulator stages (TS: 256 of the 512 columns hold the queries)
    constexpr uint32_t ACC_COL0 = TS ? 256u : 0u;
    constexpr uint32_t PIECE_COLS = (KIND == KIND_TF32X3) ? 128u : 64u;   // TMEM columns per query piece (32-bit words per row)

    extern __shared__ __align__(1024) uint8_t smem[];  // SWIZZLE_128B tiles need 1024-byte aligned bases
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if ((smem_u32(smem) & 1023u) != 0) __trap();

    uint8_t* s_q = smem;                                                         // [NA][nslab] slabs
    uint8_t* s_x = TS ? smem : s_q + static_cast<size_t>(NA) * p.nslab * SLAB_TILE;  // [n_stages][NB] slabs
    uint8_t* s_tail = s_x + static_cast<size_t>(p.n_stages) * NB * SLAB_TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_tail);
    uint64_t* bar_full = bars;                         // [n_stages]
    uint64_t* bar_empty = bars + p.n_stages;           // [n_stages]
    uint64_t* bar_q = bars + 2 * p.n_stages;           // [1]
    uint64_t* bar_tfull = bar_q + 1;                   // [ACC_STAGES]
    uint64_t* bar_tempty = bar_tfull + ACC_STAGES;     // [ACC_STAGES]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_tempty + ACC_STAGES);

    const uint32_t q0 = blockIdx.x * BM;
    const uint64_t r_begin = static_cast<uint64_t>(blockIdx.y) * p.rows_per_split;
    const uint64_t r_end = min(static_cast<uint64_t>(p.n_pad), r_begin + p.rows_per_split);
    const uint32_t n_tiles = (r_begin < r_end) ? static_cast<uint32_t>((r_end - r_begin + BN - 1) / BN) : 0;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.n_stages; s++) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }
        mbar_init(bar_q, TS ? EPI_THREADS : 1);
        for (int a = 0; a < NACC; a++) { mbar_init(bar_tfull + a, 1); mbar_init(bar_tempty + a, EPI_THREADS); }
        fence_barrier_init();
        fence_proxy_async();
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_x);
    }
    if (warp == 1) tmem_alloc(s_tmem, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            if (!TS) {
                mbar_expect_tx(bar_q, p.a_pieces * p.nslab * SLAB_TILE);
                for (uint32_t a = 0; a < p.a_pieces; a++)
                    for (uint32_t s = 0; s < p.nslab; s++)
                        tma_load_2d(smem_u32(s_q + (static_cast<size_t>(a) * p.nslab + s) * SLAB_TILE), &tm_q, bar_q, s * SLAB_ELEMS,
                                    a * p.nq_pad + q0);
            }
            uint32_t it = 0;
            long long w_prod = 0;
            for (uint32_t t = 0; t < n_tiles; t++) {
                const uint32_t row0 = static_cast<uint32_t>(r_begin) + t * BN;
                for (uint32_t s = 0; s < p.nslab; s++, it++) {
                    const uint32_t stage = it % p.n_stages, ph = (it / p.n_stages) & 1u;
                    mbar_wait_timed(bar_empty + stage, ph ^ 1u, w_prod);
                    mbar_expect_tx(bar_full + stage, NB * SLAB_TILE);
                    for (int b = 0; b < NB; b++)
                        tma_load_2d(smem_u32(s_x + (static_cast<size_t>(stage) * NB + b) * SLAB_TILE), &tm_x, bar_full + stage, s * SLAB_ELEMS,
                                    b * p.n_pad + row0);
                }
            }
            if (p.dbg_cycles && blockIdx.x == 0 && blockIdx.y == 0) p.dbg_cycles[1] = w_prod;
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        // The whole warp walks the pipeline converged (so every address / descriptor is warp-uniform and lives in
        // uniform registers); only the tcgen05 instructions themselves are issued by one elected lane.
        constexpr uint32_t idesc = make_idesc(KIND);
        constexpr uint32_t SLAB_DESC = SLAB_TILE >> 4;    // descriptor start-address units (16 B) per slab
        mbar_wait(bar_q, 0);
        tc_fence_after();
        const uint64_t q_desc0 = make_smem_desc(smem_u32(s_q));
        const uint64_t x_desc0 = make_smem_desc(smem_u32(s_x));
        uint32_t it = 0;
        long long w_full = 0, w_tempty = 0;
        const long long t_start = clock64();
        for (uint32_t t = 0; t < n_tiles; t++) {
            const uint32_t acc = t % NACC, aph = (t / NACC) & 1u;
            mbar_wait_timed(bar_tempty + acc, aph ^ 1u, w_tempty);
            tc_fence_after();
            const uint32_t tmem_c = tmem_base + ACC_COL0 + acc * BN;
            for (uint32_t s = 0; s < p.nslab; s++, it++) {
                const uint32_t stage = it % p.n_stages, ph = (it / p.n_stages) & 1u;
                mbar_wait_timed(bar_full + stage, ph, w_full);
                tc_fence_after();
                const uint64_t xd = x_desc0 + static_cast<uint64_t>(stage * NB) * SLAB_DESC;
                const uint64_t qd = q_desc0 + static_cast<uint64_t>(s) * SLAB_DESC;
                const uint64_t q_piece = static_cast<uint64_t>(p.nslab) * SLAB_DESC;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < KSTEPS; k++) {
                        const uint32_t first = (s | static_cast<uint32_t>(k)) != 0 ? 1u : 0u;   // 0 only for the tile's first MMA
                        if (TS && KIND == KIND_TF32X3) {
                            // queries in TMEM: hi at columns [0, 128), lo at [128, 256); 8 columns (8 tf32) per K step
                            const uint32_t a_hi = tmem_base + s * 32 + k * 8, a_lo = a_hi + PIECE_COLS;
                            umma_ts<KIND>(tmem_c, a_hi, xd + 2 * k, idesc, first);
                            umma_ts<KIND>(tmem_c, a_lo, xd + 2 * k, idesc, 1u);
                            umma_ts<KIND>(tmem_c, a_hi, xd + SLAB_DESC + 2 * k, idesc, 1u);
                        } else if (TS) {
                            // bf16 query terms q0, q1, q2 in TMEM at columns [0,64), [64,128), [128,192); 8 columns (16 bf16) per K step
                            const uint32_t a0 = tmem_base + s * 32 + k * 8;
                            umma_ts<KIND>(tmem_c, a0, xd + 2 * k, idesc, first);
                            if (p.a_pieces > 1) {
                                umma_ts<KIND>(tmem_c, a0 + PIECE_COLS, xd + 2 * k, idesc, 1u);
                                umma_ts<KIND>(tmem_c, a0 + 2 * PIECE_COLS, xd + 2 * k, idesc, 1u);
                            }
                        } else if (KIND == KIND_TF32X3) {
                            // s = Qhi.Xhi + Qlo.Xhi + Qhi.Xlo   (the lo.lo term is below 2^-22 relative)
                            umma<KIND>(tmem_c, qd + 2 * k, xd + 2 * k, idesc, first);
                            umma<KIND>(tmem_c, qd + q_piece + 2 * k, xd + 2 * k, idesc, 1u);
                            umma<KIND>(tmem_c, qd + 2 * k, xd + SLAB_DESC + 2 * k, idesc, 1u);
                        } else {
                            umma<KIND>(tmem_c, qd + 2 * k, xd + 2 * k, idesc, first);
                            if (p.a_pieces > 1) {
                                umma<KIND>(tmem_c, qd + q_piece + 2 * k, xd + 2 * k, idesc, 1u);
                                umma<KIND>(tmem_c, qd + 2 * q_piece + 2 * k, xd + 2 * k, idesc, 1u);
                            }
                        }
                    }
                    umma_commit(bar_empty + stage);                       // slab consumed once the MMAs above retire
                    if (s + 1 == p.nslab) umma_commit(bar_tfull + acc);   // accumulator ready for the epilogue
                }
                __syncwarp();
            }
        }
        if (p.dbg_cycles && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) {
            p.dbg_cycles[0] = clock64() - t_start;
            p.dbg_cycles[2] = w_full;
            p.dbg_cycles[3] = w_tempty;
            p.dbg_cycles[6] = n_tiles;
        }
        __syncwarp();
    } else {
        // ===================================================================== epilogue (4 warps, thread = query row)
        const uint32_t quarter = warp & 3u;                 // TMEM lane quarter this warp may access
        const uint32_t row_in_tile = quarter * 32 + lane;   // query row inside the tile
        const uint32_t half = (warp - 2) >> 2;              // which 64-column half of every tile this warp scans
        const uint32_t e = threadIdx.x - 64;                // 0..255; the first 128 stage aux
        TopList<KP> top;
        top.init();
        float scratch[64];
        long long w_tfull = 0, w_slow = 0;
        // Shared threshold: the k'-th best value any CTA of this query has seen so far (monotone, atomicMin on the
        // order-preserving integer image).  A value that does not beat it cannot be in the merged top-k', whichever
        // split holds it, so every split prunes with the tightest bound known anywhere.  Stale reads are only looser.
        if (TS) {
            // Stage this CTA's 128 queries into TMEM once: query pieces are dealt to the two warps of each lane quarter
            // (even pieces to half 0, odd to half 1); thread = query row = TMEM lane, 32 columns (128 B of the row) per
            // tcgen05.st.  16-bit operands sit two per column, low half = even k, exactly as in memory.
            const uint32_t row_words = p.kp * ELEM / 4;
            for (uint32_t pc = half; pc < p.a_pieces; pc += 2) {
                const uint32_t* src = reinterpret_cast<const uint32_t*>(p.q_op) + (static_cast<size_t>(pc) * p.nq_pad + q0 + row_in_tile) * row_words;
                const uint32_t tq = tmem_base + ((quarter * 32u) << 16) + pc * PIECE_COLS;
                for (uint32_t c = 0; c < row_words; c += 32) {
                    uint32_t w[32];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const uint4 x = __ldg(reinterpret_cast<const uint4*>(src + c) + j);
                        w[4 * j] = x.x; w[4 * j + 1] = x.y; w[4 * j + 2] = x.z; w[4 * j + 3] = x.w;
                    }
                    tmem_st32(tq + c, w);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_q);
        }
        uint32_t* gtau_ptr = p.gtau + q0 + row_in_tile;
        uint32_t g_next = *reinterpret_cast<volatile uint32_t*>(gtau_ptr);
        // Per-column constants of this warp's 64-column half: lane l keeps columns l and l + 32 in registers
        // (coalesced load, fetched one tile ahead) and the warp broadcasts them with shuffles -- no shared memory,
        // no barrier between the epilogue warps.
        const float* aux_half = p.aux + r_begin + half * 64 + lane;
        float aux_lo_next = 0.f, aux_hi_next = 0.f;
        if (n_tiles > 0) { aux_lo_next = __ldg(aux_half); aux_hi_next = __ldg(aux_half + 32); }
        for (uint32_t t = 0; t < n_tiles; t++) {
            const uint32_t acc = t % NACC, aph = (t / NACC) & 1u;
            const uint32_t row0 = static_cast<uint32_t>(r_begin) + t * BN;
            const uint32_t g_bits = g_next;
            const float aux_lo = aux_lo_next, aux_hi = aux_hi_next;
            if (t + 1 < n_tiles) {   // prefetch for the next tile: latency hidden behind this tile's work
                aux_lo_next = __ldg(aux_half + static_cast<size_t>(t + 1) * BN);
                aux_hi_next = __ldg(aux_half + static_cast<size_t>(t + 1) * BN + 32);
                g_next = *reinterpret_cast<volatile uint32_t*>(gtau_ptr);
            }
            mbar_wait_timed(bar_tfull + acc, aph, w_tfull);
            tc_fence_after();
            const float g_tau = ordered_to_f32(g_bits);       // NaN (all-ones init) is ignored by fminf
            float tau = fminf(top.tau(), g_tau);
            const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + ACC_COL0 + acc * BN;
            {
                const int c = static_cast<int>(half);
                uint32_t r[64];
                tmem_ld32(taddr + c * 64, r);
                tmem_ld32(taddr + c * 64 + 32, r + 32);
                tmem_ld_wait();
                float v[64];
                float gm[8];  // minima of the 8 groups of 8 columns
#pragma unroll
                for (int g = 0; g < 8; g++) {
                    float mg = INFINITY;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int col = g * 8 + j;   // compile-time after unrolling
                        const float cst = __shfl_sync(0xFFFFFFFFu, col < 32 ? aux_lo : aux_hi, col & 31);
                        const float sdot = __uint_as_float(r[col]);
                        v[col] = (MET == MET_L2) ? fmaf(sdot, -2.0f, cst) : sdot * cst;
                        mg = fminf(mg, v[g * 8 + j]);
                    }
                    gm[g] = mg;
                }
                const float m = fminf(fminf(fminf(gm[0], gm[1]), fminf(gm[2], gm[3])), fminf(fminf(gm[4], gm[5]), fminf(gm[6], gm[7])));
                if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && t == 0) {
#pragma unroll
                    for (int j = 0; j < 64; j++) p.dbg[row_in_tile * BN + c * 64 + j] = v[j];
                }
                if (m < tau) {
                    const long long ts0 = clock64();
                    // Rare path, cost proportional to the number of groups that really hold a candidate: stash the
                    // 64 values once (independent stores), then visit only the groups whose minimum beats tau.
#pragma unroll
                    for (int j = 0; j < 64; j++) scratch[j] = v[j];
                    uint32_t gmask = 0;
#pragma unroll
                    for (int g = 0; g < 8; g++) gmask |= (gm[g] < tau) ? (1u << g) : 0u;
                    while (gmask) {
                        const int g = __ffs(gmask) - 1;
                        gmask &= gmask - 1;
                        float s8[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) s8[j] = scratch[g * 8 + j];
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            if (s8[j] < tau) {
                                top.insert(s8[j], row0 + c * 64 + g * 8 + j);
                                tau = fminf(tau, top.tau());
                            }
                        }
                    }
                    w_slow += clock64() - ts0;
                }
            }
            tc_fence_before();
            mbar_arrive(bar_tempty + acc);
            if (top.tau() < g_tau || (g_tau != g_tau && top.tau() < INFINITY)) atomicMin(gtau_ptr, f32_to_ordered(top.tau()));
        }
        if (p.dbg_cycles && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64) { p.dbg_cycles[4] = w_tfull; p.dbg_cycles[5] = w_slow; }
        const uint64_t q = static_cast<uint64_t>(q0) + row_in_tile;
        if (q < p.nq) {
            uint64_t* out = p.part_keys + (q * (2 * p.n_splits) + 2 * blockIdx.y + half) * KP;
#pragma unroll
            for (int j = 0; j < KP; j++) out[j] = (top.i[j] == IDX_INVALID) ? KEY_SENTINEL : make_key(top.v[j], top.i[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ----------------------------------------------------------------------------------------------- operand preparation
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// f32 rows (pitch ld_src floats) -> stacked [2][rows_pad][kp] tf32 hi / lo, zero padded.
__global__ void split_tf32_kernel(const float* __restrict__ src, uint32_t ld_src, uint32_t dim, uint64_t rows, uint64_t rows_pad, uint32_t kp,
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
// f32 queries -> stacked [3][rows_pad][kp] bf16 terms q0 + q1 + q2 (each RNE), zero padded.
__global__ void split_bf16x3_kernel(const float* __restrict__ src, uint32_t ld_src, uint32_t dim, uint64_t rows, uint64_t rows_pad, uint32_t kp,
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
__global__ void pad_bf16_kernel(const uint16_t* __restrict__ src, uint32_t ld_src, uint32_t dim, uint64_t rows, uint64_t rows_pad, uint32_t kp,
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
__global__ void aux_kernel(const uint8_t* __restrict__ rows, uint32_t row_bytes, int is_bf16, uint32_t dim, const float* __restrict__ norms,
                           uint64_t n, uint64_t n_pad_total, float* __restrict__ aux) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n_pad_total) return;
    float o = INFINITY;
    if (i < n) {
        if (norms) {
            o = -1.0f / norms[i];
        } else {
            float s = 0.f;
            const uint8_t* r = rows + i * row_bytes;
            for (uint32_t e = 0; e < dim; e++) {
                const float x = is_bf16 ? bf16_bits_to_f32(reinterpret_cast<const uint16_t*>(r)[e]) : reinterpret_cast<const float*>(r)[e];
                s = fmaf(x, x, s);
            }
            o = s;
        }
    }
    aux[i] = o;
}

// ----------------------------------------------------------------------------------------------- exact re-rank + merge
struct RerankParams {
    const uint64_t* part_keys;  // [nq][parts][kp] approximate keys
    uint32_t parts, kp, k_eff, k_out, nsort;
    uint64_t nq;
    const uint8_t* rows;  // index rows in the index dtype
    uint32_t row_bytes;
    const float* row_norms;
    const uint8_t* queries;  // prepared queries (f32 padded rows, or bf16 rows for self queries)
    uint32_t q_bytes;
    uint32_t dim;
    int bf16_self;
    uint64_t id_base;
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
    __shared__ uint64_t exact[64];
    const uint64_t q = blockIdx.x;
    const uint32_t total = p.parts * p.kp;
    const uint64_t* src = p.part_keys + q * total;
    for (uint32_t i = threadIdx.x; i < p.nsort; i += blockDim.x) keys[i] = (i < total) ? src[i] : KEY_SENTINEL;
    __syncthreads();
    bitonic_sort_keys<true>(keys, p.nsort, threadIdx.x, blockDim.x);
    if (threadIdx.x < 64) {
        uint64_t ek = KEY_SENTINEL;
        if (threadIdx.x < p.kp) {
            const uint64_t key = keys[threadIdx.x];
            const uint32_t idx = key_idx(key);
            if (idx != IDX_INVALID) {
                const uint8_t* row = p.rows + static_cast<uint64_t>(idx) * p.row_bytes;
                const uint8_t* qv = p.queries + q * p.q_bytes;
                float raw[1];
                accumulate_fp<(RT == 0) ? 4 : 2, (QT == QT_F32) ? 4 : 2, MET == MET_L2, 1>(row, qv, p.q_bytes, p.dim, raw);
                float qn = 1.0f, xn = 1.0f;
                if (MET == MET_COS) {
                    qn = seq_norm<(QT == QT_F32) ? 4 : 2>(qv, p.dim);
                    if (p.bf16_self) qn = round_to_bf16(qn);
                    xn = p.row_norms[idx];
                }
                ek = make_key(finish_fp<MET>(raw[0], qn, xn), idx);
            }
        }
        exact[threadIdx.x] = ek;
    }
    __syncthreads();
    if (threadIdx.x < 32) bitonic_sort_keys<false>(exact, 64, threadIdx.x, 32);
    __syncthreads();
    uint32_t valid = 0;
    for (uint32_t j = threadIdx.x; j < p.k_out; j += blockDim.x) {
        uint64_t key = (j < p.k_eff && j < 64) ? exact[j] : KEY_SENTINEL;
        uint64_t id = 0xFFFFFFFFFFFFFFFFull;
        float d = INFINITY;
        if (key_idx(key) != IDX_INVALID) {
            id = static_cast<uint64_t>(key_idx(key)) + p.id_base;
            d = key_dist(key);
            valid++;
        }
        p.out_ids[q * p.k_out + j] = id;
        if (p.out_dist) p.out_dist[q * p.k_out + j] = d;
    }
    if (p.out_counts) {
        __shared__ uint32_t s_cnt;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        if (valid) atomicAdd(&s_cnt, valid);
        __syncthreads();
        if (threadIdx.x == 0) p.out_counts[q] = s_cnt;
    }
}

}  // namespace tc

// =============================================================================================== host side
struct TcState {
    int kind = -1;
    uint32_t kp_elems = 0;  // padded K in elements
    uint32_t nslab = 0;
    uint32_t n_pad = 0;
    void* d_x = nullptr;      // stacked database operand (nullptr: the index rows themselves are used)
    float* d_aux = nullptr;   // [n_pad + BN]
    CUtensorMap tm_x;
    DevBuf q_op, part, dbg, gtau, dbgc;
    uint64_t bytes = 0;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

// 2-D row-major [rows][kp] tensor, box = {128 bytes of K, 128 rows}, SWIZZLE_128B.
static int make_tmap(CUtensorMap* tm, void* base, uint64_t rows, uint32_t kp_elems, int elem_bytes) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_last_error("cuTensorMapEncodeTiled is unavailable (driver too old?)"); return ANNB_ERR_CUDA; }
    cuuint64_t dims[2] = {kp_elems, rows};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(kp_elems) * elem_bytes};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(tc::SLAB_BYTES / elem_bytes), static_cast<cuuint32_t>(tc::BM)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r))); return ANNB_ERR_CUDA; }
    return ANNB_OK;
}

static inline uint32_t blocks_for(uint64_t work) { return static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>((work + 255) / 256, 148 * 16))); }

int tc_flat_prepare(annb_index* ix) {
    if (ix->is_ivf || ix->dtype == ANNB_SQ8) return ANNB_OK;
    const int kind = ix->dtype == ANNB_F32 ? tc::KIND_TF32X3 : tc::KIND_BF16;
    const uint32_t elem = kind == tc::KIND_TF32X3 ? 4 : 2;
    const uint32_t slab_elems = tc::SLAB_BYTES / elem;
    const uint32_t kp = round_up(ix->dim, slab_elems);
    if (kp * elem > 512) return ANNB_OK;  // query tile would not fit in shared memory: the exact CUDA-core path serves this index
    if (ix->n < 4096) return ANNB_OK;     // tiny indices stay on the exact CUDA-core path
    TcState* st = new TcState();
    ix->tc = st;
    st->kind = kind;
    st->kp_elems = kp;
    st->nslab = kp / slab_elems;
    st->n_pad = static_cast<uint32_t>(round_up<uint64_t>(ix->n, tc::BN));
    cudaStream_t s = ix->stream;
    const uint64_t aux_rows = static_cast<uint64_t>(st->n_pad) + tc::BN;
    {
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&st->d_aux), aux_rows * sizeof(float));
        if (e != cudaSuccess) { (void)cudaGetLastError(); set_last_error(std::string("cudaMalloc tc aux: ") + cudaGetErrorString(e)); return ANNB_ERR_OUT_OF_MEMORY; }
        st->bytes += aux_rows * sizeof(float);
    }
    tc::aux_kernel<<<static_cast<uint32_t>((aux_rows + 127) / 128), 128, 0, s>>>(ix->d_rows, ix->row_bytes, kind == tc::KIND_BF16, ix->dim,
                                                                                ix->metric == ANNB_COSINE ? ix->d_norms : nullptr, ix->n, aux_rows,
                                                                                st->d_aux);
    ANNB_CUDA_CHECK(cudaGetLastError());
    void* xbase = nullptr;
    uint64_t xrows = 0;
    if (kind == tc::KIND_TF32X3) {
        const uint64_t bytes = 2ull * st->n_pad * kp * 4;
        cudaError_t e = cudaMalloc(&st->d_x, bytes);
        if (e != cudaSuccess) { (void)cudaGetLastError(); set_last_error(std::string("cudaMalloc tc operand: ") + cudaGetErrorString(e)); return ANNB_ERR_OUT_OF_MEMORY; }
        st->bytes += bytes;
        tc::split_tf32_kernel<<<blocks_for(static_cast<uint64_t>(st->n_pad) * kp), 256, 0, s>>>(reinterpret_cast<const float*>(ix->d_rows), ix->row_bytes / 4,
                                                                                                ix->dim, ix->n, st->n_pad, kp, static_cast<float*>(st->d_x));
        ANNB_CUDA_CHECK(cudaGetLastError());
        xbase = st->d_x;
        xrows = 2ull * st->n_pad;
    } else {
        const uint64_t bytes = static_cast<uint64_t>(st->n_pad) * kp * 2;
        cudaError_t e = cudaMalloc(&st->d_x, bytes);
        if (e != cudaSuccess) { (void)cudaGetLastError(); set_last_error(std::string("cudaMalloc tc operand: ") + cudaGetErrorString(e)); return ANNB_ERR_OUT_OF_MEMORY; }
        st->bytes += bytes;
        tc::pad_bf16_kernel<<<blocks_for(static_cast<uint64_t>(st->n_pad) * kp), 256, 0, s>>>(reinterpret_cast<const uint16_t*>(ix->d_rows), ix->row_bytes / 2,
                                                                                              ix->dim, ix->n, st->n_pad, kp, static_cast<uint16_t*>(st->d_x));
        ANNB_CUDA_CHECK(cudaGetLastError());
        xbase = st->d_x;
        xrows = st->n_pad;
    }
    ANNB_TRY(make_tmap(&st->tm_x, xbase, xrows, kp, elem));
    ix->device_bytes += st->bytes;
    return ANNB_OK;
}

void tc_destroy(annb_index* ix) {
    if (!ix->tc) return;
    cudaFree(ix->tc->d_x);
    cudaFree(ix->tc->d_aux);
    ix->tc->q_op.release();
    ix->tc->part.release();
    ix->tc->dbg.release();
    ix->tc->gtau.release();
    ix->tc->dbgc.release();
    delete ix->tc;
    ix->tc = nullptr;
}

static uint32_t pick_kprime(const annb_index* ix, uint32_t k_eff) {
    if (ix->opt_tc_candidates == 16 || ix->opt_tc_candidates == 32) return std::max<uint32_t>(ix->opt_tc_candidates, k_eff <= 16 ? 16 : 32);
    return k_eff <= 10 ? 16 : 32;
}

bool tc_flat_supported(const annb_index* ix, int qt, uint32_t k_eff) {
    if (!ix->tc) return false;
    if (k_eff > 24) return false;
    if (ix->dtype == ANNB_F32) return qt == QT_F32;
    if (ix->dtype == ANNB_BF16) return qt == QT_F32 || qt == QT_BF16;
    return false;
}

// Database splits per query tile: fill whole waves of 148 CTAs.
static uint32_t pick_splits(uint64_t q_tiles, uint64_t n_tiles_db, int requested) {
    if (requested > 0) return static_cast<uint32_t>(std::min<uint64_t>(requested, n_tiles_db));
    const uint64_t sms = 148;
    uint32_t best = 1;
    double best_eff = -1.0;
    const uint64_t smax = std::min<uint64_t>(n_tiles_db, 64);
    for (uint64_t s = 1; s <= smax; s++) {
        const uint64_t tiles_per = (n_tiles_db + s - 1) / s;
        const uint64_t splits = (n_tiles_db + tiles_per - 1) / tiles_per;
        const uint64_t ctas = q_tiles * splits;
        const uint64_t waves = (ctas + sms - 1) / sms;
        // work per CTA shrinks with more splits; fixed per-CTA cost ~ 8 tiles' worth (query load, pipeline fill, warm-up of the k' list)
        const double eff = (static_cast<double>(ctas) / static_cast<double>(waves * sms)) * (static_cast<double>(tiles_per) / static_cast<double>(tiles_per + 8));
        if (eff > best_eff + 1e-9) { best_eff = eff; best = static_cast<uint32_t>(splits); }
    }
    return best;
}

template <int KIND, int KP, int MET, bool TS>
static int launch_tc(const CUtensorMap& tmq, const CUtensorMap& tmx, const tc::Params& p, dim3 grid, size_t smem, cudaStream_t s) {
    auto kern = tc::flat_tc_kernel<KIND, KP, MET, TS>;
    ANNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, tc::NUM_THREADS, smem, s>>>(tmq, tmx, p);
    ANNB_CUDA_CHECK(cudaGetLastError());
    return ANNB_OK;
}

template <int RT, int QT, int MET>
static int launch_rerank(const tc::RerankParams& r, cudaStream_t s) {
    auto kern = tc::rerank_kernel<RT, QT, MET>;
    const size_t smem = static_cast<size_t>(r.nsort) * 8;
    ANNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<static_cast<uint32_t>(r.nq), 128, smem, s>>>(r);
    ANNB_CUDA_CHECK(cudaGetLastError());
    return ANNB_OK;
}

int tc_flat_search(annb_index* ix, const uint8_t* d_q, uint32_t q_bytes, int qt, int bf16_self, uint64_t nq, uint32_t k_eff, uint32_t k_out,
                   uint64_t* d_ids, float* d_dist, uint32_t* d_cnt, cudaStream_t s) {
    TcState* st = ix->tc;
    const int kind = st->kind;
    const uint32_t elem = kind == tc::KIND_TF32X3 ? 4 : 2;
    const uint32_t kp = st->kp_elems;
    const uint32_t kprime = pick_kprime(ix, k_eff);
    const uint32_t nq_pad = static_cast<uint32_t>(round_up<uint64_t>(nq, tc::BM));
    const uint32_t na = kind == tc::KIND_TF32X3 ? 2 : (qt == QT_BF16 ? 1 : 3);

    // ---- query operand: stacked pieces, zero padded ----
    ANNB_TRY(st->q_op.ensure(static_cast<uint64_t>(kind == tc::KIND_TF32X3 ? 2 : 3) * nq_pad * kp * elem));
    if (kind == tc::KIND_TF32X3) {
        tc::split_tf32_kernel<<<blocks_for(static_cast<uint64_t>(nq_pad) * kp), 256, 0, s>>>(reinterpret_cast<const float*>(d_q), q_bytes / 4, ix->dim, nq, nq_pad, kp,
                                                                                             st->q_op.as<float>());
    } else if (qt == QT_F32) {
        tc::split_bf16x3_kernel<<<blocks_for(static_cast<uint64_t>(nq_pad) * kp), 256, 0, s>>>(reinterpret_cast<const float*>(d_q), q_bytes / 4, ix->dim, nq, nq_pad, kp,
                                                                                               st->q_op.as<__nv_bfloat16>());
    } else {
        tc::pad_bf16_kernel<<<blocks_for(static_cast<uint64_t>(nq_pad) * kp), 256, 0, s>>>(reinterpret_cast<const uint16_t*>(d_q), q_bytes / 2, ix->dim, nq, nq_pad, kp,
                                                                                           st->q_op.as<uint16_t>());
    }
    ANNB_CUDA_CHECK(cudaGetLastError());
    ix->stat_launches++;
    CUtensorMap tmq;
    ANNB_TRY(make_tmap(&tmq, st->q_op.p, static_cast<uint64_t>(na) * nq_pad, kp, elem));

    // ---- geometry ----
    const uint64_t q_tiles = nq_pad / tc::BM;
    const uint64_t db_tiles = st->n_pad / tc::BN;
    const uint32_t splits_req = pick_splits(q_tiles, db_tiles, ix->opt_db_splits);
    const uint64_t tiles_per = (db_tiles + splits_req - 1) / splits_req;
    const uint32_t splits = static_cast<uint32_t>((db_tiles + tiles_per - 1) / tiles_per);
    const uint32_t nb = kind == tc::KIND_TF32X3 ? 2 : 1;
    const bool ts = ix->opt_tc_ts != 0;   // query operand resident in TMEM (TS-mode MMA)
    const size_t q_smem = ts ? 0 : static_cast<size_t>(kind == tc::KIND_TF32X3 ? 2 : 3) * st->nslab * tc::SLAB_TILE;
    const size_t fixed = 256 /*barriers*/;
    const size_t budget = 227 * 1024;
    if (q_smem + fixed + nb * tc::SLAB_TILE > budget) { set_last_error("tensor path: query tile too large for shared memory"); return ANNB_ERR_UNSUPPORTED; }
    uint32_t stages = static_cast<uint32_t>((budget - q_smem - fixed) / (nb * tc::SLAB_TILE));
    stages = std::min<uint32_t>(stages, 8);
    const size_t smem = q_smem + static_cast<size_t>(stages) * nb * tc::SLAB_TILE + fixed;

    ANNB_TRY(st->part.ensure(nq * 2 * splits * static_cast<uint64_t>(kprime) * 8));
    ANNB_TRY(st->gtau.ensure(static_cast<uint64_t>(nq_pad) * 4));
    ANNB_CUDA_CHECK(cudaMemsetAsync(st->gtau.p, 0xFF, static_cast<uint64_t>(nq_pad) * 4, s));
    tc::Params p{};
    p.nq = nq; p.n_rows = ix->n; p.nq_pad = nq_pad; p.n_pad = st->n_pad; p.nslab = st->nslab; p.n_stages = stages;
    p.n_splits = splits; p.rows_per_split = tiles_per * tc::BN; p.a_pieces = na; p.aux = st->d_aux;
    p.part_keys = st->part.as<uint64_t>(); p.dbg = st->dbg.as<float>(); p.gtau = st->gtau.as<uint32_t>(); p.q_op = st->q_op.as<void>(); p.kp = kp; p.dbg_cycles = st->dbgc.as<unsigned long long>();
    {
        dim3 grid(static_cast<uint32_t>(q_tiles), splits);
        // timed as the dominant kernel of the flat path
        cudaEvent_t ea = nullptr, eb = nullptr;
        if (ix->opt_time_kernels && cudaEventCreate(&ea) == cudaSuccess && cudaEventCreate(&eb) == cudaSuccess) cudaEventRecord(ea, s);
        int rc;
        const bool l2 = ix->metric == ANNB_L2;
#define ANNB_TC_LAUNCH(KIND_, KP_, TS_) (l2 ? launch_tc<KIND_, KP_, MET_L2, TS_>(tmq, st->tm_x, p, grid, smem, s) : launch_tc<KIND_, KP_, MET_COS, TS_>(tmq, st->tm_x, p, grid, smem, s))
        if (kind == tc::KIND_TF32X3 && ts) rc = kprime == 16 ? ANNB_TC_LAUNCH(tc::KIND_TF32X3, 16, true) : ANNB_TC_LAUNCH(tc::KIND_TF32X3, 32, true);
        else if (kind == tc::KIND_TF32X3) rc = kprime == 16 ? ANNB_TC_LAUNCH(tc::KIND_TF32X3, 16, false) : ANNB_TC_LAUNCH(tc::KIND_TF32X3, 32, false);
        else if (ts) rc = kprime == 16 ? ANNB_TC_LAUNCH(tc::KIND_BF16, 16, true) : ANNB_TC_LAUNCH(tc::KIND_BF16, 32, true);
        else rc = kprime == 16 ? ANNB_TC_LAUNCH(tc::KIND_BF16, 16, false) : ANNB_TC_LAUNCH(tc::KIND_BF16, 32, false);
#undef ANNB_TC_LAUNCH
        if (ea && eb) { cudaEventRecord(eb, s); ix->timed.emplace_back(ea, eb); }
        ANNB_TRY(rc);
        ix->stat_launches++;
    }
    // ---- exact re-rank + merge ----
    tc::RerankParams r{};
    r.part_keys = st->part.as<uint64_t>(); r.parts = 2 * splits; r.kp = kprime; r.k_eff = k_eff; r.k_out = k_out;
    r.nsort = next_pow2(std::max(2 * splits * kprime, 64u));
    r.nq = nq; r.rows = ix->d_rows; r.row_bytes = ix->row_bytes; r.row_norms = ix->d_norms; r.queries = d_q; r.q_bytes = q_bytes; r.dim = ix->dim;
    r.bf16_self = bf16_self; r.id_base = ix->id_base; r.out_ids = d_ids; r.out_dist = d_dist; r.out_counts = d_cnt;
    const bool cos = ix->metric == ANNB_COSINE;
    int rc;
    if (ix->dtype == ANNB_F32) rc = cos ? launch_rerank<0, QT_F32, MET_COS>(r, s) : launch_rerank<0, QT_F32, MET_L2>(r, s);
    else if (qt == QT_F32) rc = cos ? launch_rerank<1, QT_F32, MET_COS>(r, s) : launch_rerank<1, QT_F32, MET_L2>(r, s);
    else rc = cos ? launch_rerank<1, QT_BF16, MET_COS>(r, s) : launch_rerank<1, QT_BF16, MET_L2>(r, s);
    ANNB_TRY(rc);
    ix->stat_launches++;
    return ANNB_OK;
}

// Debug hook used by tests: allocate / fetch the 128 x 128 tile dump of CTA (0, 0).
int tc_debug_enable(annb_index* ix, bool on) {
    if (!ix->tc) return ANNB_ERR_UNSUPPORTED;
    if (on) { ANNB_TRY(ix->tc->dbgc.ensure(64)); cudaMemset(ix->tc->dbgc.p, 0, 64); return ix->tc->dbg.ensure(tc::BM * tc::BN * sizeof(float)); }
    ix->tc->dbg.release();
    ix->tc->dbgc.release();
    return ANNB_OK;
}
int tc_debug_fetch(annb_index* ix, float* host_out) {
    if (!ix->tc || !ix->tc->dbg.p) return ANNB_ERR_UNSUPPORTED;
    ANNB_CUDA_CHECK(cudaMemcpy(host_out, ix->tc->dbg.p, tc::BM * tc::BN * sizeof(float), cudaMemcpyDeviceToHost));
    // the first 8 floats' worth of the dump are followed out-of-band by the cycle counters (see tc_debug_cycles)
    return ANNB_OK;
}
int tc_debug_cycles(annb_index* ix, unsigned long long* host_out8) {
    if (!ix->tc || !ix->tc->dbgc.p) return ANNB_ERR_UNSUPPORTED;
    ANNB_CUDA_CHECK(cudaMemcpy(host_out8, ix->tc->dbgc.p, 64, cudaMemcpyDeviceToHost));
    return ANNB_OK;
}

}  // namespace annb
