// flat_tc.cu -- tensor-core flat search: tcgen05.mma into TMEM, operands fed by TMA, top-k' selection fused
// into the TMEM epilogue (the [nq x n] distance matrix never exists), then an exact re-rank of the k'
// survivors in the reference's own floating-point order (refdist.cuh).
//
// Replaces, for the flat f32 / bf16 / SQ8 indices, the reference's distance-matrix kernels + separate top-k pass
// (src/gpu/dist_gpu.rs:79-488 euclidean/cosine_tiled{,_reg}; :553-613 extract_topk; src/gpu/topk_gpu.rs:992-1237), and with
// DENSE + coarse_select_gm_kernel / coarse_select_kernel the centroid ranking and probe expansion of the IVF query
// (src/cpu/ivf.rs:349-365, src/utils/k_means_utils.rs:3007-3029).
//
// Kernel shape (one CTA = 128 queries x one database split, 320 threads, 1 CTA / SM):
//   warp 0      TMA producer: a ring of database K-slabs (128 rows x 128 B, SWIZZLE_128B)
//   warp 1      TMEM allocator + tcgen05.mma issuer (converged warp, elect.sync), accumulator ring in TMEM
//   warps 2..9  epilogue, two per TMEM lane quarter (one 64-column half each): stage the query tile into TMEM once (TS
//               mode), then per tile tcgen05.ld 64 columns, v = fma(s, -2, |x|^2) or s * (-1/|x|) with the row constants
//               from a warp-private shared-memory row, min-tree + threshold test; a value that beats the threshold is
//               inserted into a register-resident sorted k' list (select_from_tile)
//   f32 index : 3xFP16 (rows scaled by powers of two, kind::f16) or 3xTF32 split precision  s = Qhi.Xhi + Qlo.Xhi + Qhi.Xlo   (hi/lo rounded to nearest)
//   bf16 index: f32 queries split into three bf16 terms, s = (q0 + q1 + q2).X  (kind::f16), X = the stored bf16 rows
//   SQ8 index : int8 codes x int8 codes, s32 accumulators (kind::i8): exact integer dots, three accumulator stages
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "flat_tc.hpp"
#include "index.hpp"
#include "tc_common.cuh"

namespace annb {

namespace tc {

// ----------------------------------------------------------------------------------------------- the kernel
// TS = the query operand lives in TMEM (columns [0, 256): hi then lo) instead of shared memory: the MMAs read only the
// database slab from shared memory (half the operand bandwidth) and the whole 227 KB becomes database ring.
// DENSE = write every selection value to p.dense instead of keeping a top-k' (the IVF centroid ranking consumes the full
// [nq x nlist] matrix of approximate values and, optionally, the minima of its aligned groups of 8; see coarse_select_gm_kernel).
__device__ __forceinline__ bool hyb_cfg(const Params& p) { return p.hybrid != 0; }

// EW = epilogue warps per TMEM lane quarter (2 or 4): every warp owns 128 / EW columns of each tile.  The kernels whose
// epilogue paces them (int8 codes, bf16 terms: few MMAs per tile) run four, so that four warps per scheduler interleave
// their load -> transform -> select chains; the MMA-bound f32 kernels keep two (and their larger register budget).
template <int KIND, int KP, int MET, bool TS, bool DENSE = false, int EW = 2>
__global__ void __launch_bounds__(64 + EW * 128, 1)
flat_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_x, const Params p) {
    constexpr bool SPLIT3 = (KIND == KIND_TF32X3 || KIND == KIND_F16X3);   // hi / lo pieces on both sides, three product terms
    constexpr int NA = SPLIT3 ? 2 : (KIND == KIND_I8 ? 1 : 3);  // stacked query pieces
    constexpr int NB = SPLIT3 ? 2 : 1;  // stacked database pieces
    constexpr int ELEM = (KIND == KIND_TF32X3) ? 4 : (KIND == KIND_I8 ? 1 : 2);
    constexpr int SLAB_ELEMS = SLAB_BYTES / ELEM;      // 32 tf32 / 64 bf16 per slab row
    constexpr int KSTEPS = 4;                          // 128 B / 32 B per UMMA K step (8 tf32 / 16 bf16)
    constexpr int NV = BN / EW;                        // columns per epilogue warp and tile
    constexpr int NG = NV / 8;
    constexpr uint32_t EPI_T = EW * 128;               // epilogue threads
    // TMEM columns per query piece (32-bit words per row): f32 128; bf16 terms 64 (rows of <= 128 elements) or 128; int8 codes 128
    const uint32_t PIECE_COLS = (KIND == KIND_TF32X3) ? (p.lo_smem ? p.kp : 128u)
                                : (KIND == KIND_F16X3 ? p.kp / 2u : (KIND == KIND_I8 ? 128u : (p.kp > 128u ? 128u : 64u)));
    // accumulator stages behind the TMEM-resident queries (TS): three when the pieces take at most 128 columns (int8 codes, one
    // or two bf16 terms of narrow rows), two when they take up to 256 (f32 hi / lo, three bf16 terms, two bf16 terms of wide rows)
    const uint32_t q_cols = (KIND == KIND_I8) ? 128u : ((SPLIT3 && p.lo_smem) ? PIECE_COLS : (KIND == KIND_F16X3 ? 2u * PIECE_COLS : (hyb_cfg(p) ? 2u : p.a_pieces) * PIECE_COLS));
    const uint32_t NACC = TS ? (q_cols <= 128u ? 3u : 2u) : static_cast<uint32_t>(ACC_STAGES);
    const uint32_t ACC_COL0 = TS ? (q_cols <= 128u ? 128u : 256u) : 0u;

    extern __shared__ __align__(1024) uint8_t smem[];  // SWIZZLE_128B tiles need 1024-byte aligned bases
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if ((smem_u32(smem) & 1023u) != 0) __trap();

    // Hybrid placement (bf16 index, f32 queries): terms q0, q1 live in TMEM, q2 stays in shared memory and is multiplied by an
    // SS-mode MMA.  The A operand of a TS-mode MMA and the epilogue's tcgen05.ld share the TMEM read path: three TMEM terms
    // read 96 KB per tile on top of the epilogue's 64 KB, which is what paces the bf16 kernel.
    const bool hyb = TS && KIND == KIND_BF16 && p.hybrid != 0;
    // f32 rows of more than 128 elements (option: of any width): only the hi piece fits TMEM beside two accumulator stages; the lo
    // piece stays in shared memory and its term Qlo.Xhi is issued as an SS-mode MMA.
    const bool lo_s = TS && SPLIT3 && p.lo_smem != 0;
    uint8_t* s_q = smem;                                                         // [NA][nslab] slabs (hybrid: [nslab], term q2; f32 lo_smem: [nslab], piece lo)
    uint8_t* s_x = TS ? ((hyb || lo_s) ? smem + static_cast<size_t>(p.nslab) * SLAB_TILE : smem)
                      : s_q + static_cast<size_t>(NA) * p.nslab * SLAB_TILE;     // [n_stages][NB] slabs
    // stream_q (SS mode only, rows too wide for a resident query tile): the query slabs travel through the ring with the database
    // slabs -- one stage = [a_pieces query slabs | NB database slabs] of the same K slab; the tile's queries are re-read from L2
    // for every database tile.
    const bool stream_q = !TS && p.stream_q != 0;
    const uint32_t q_slabs = stream_q ? p.a_pieces : 0u;          // query slabs per ring stage
    const uint32_t stage_slabs = q_slabs + NB;
    if (stream_q) s_x = smem;
    uint8_t* s_tail = s_x + static_cast<size_t>(p.n_stages) * stage_slabs * SLAB_TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_tail);
    uint64_t* bar_full = bars;                         // [n_stages]
    uint64_t* bar_empty = bars + p.n_stages;           // [n_stages]
    uint64_t* bar_q = bars + 2 * p.n_stages;           // [1]
    uint64_t* bar_tfull = bar_q + 1;                   // [ACC_STAGES]
    uint64_t* bar_tempty = bar_tfull + ACC_STAGES;     // [ACC_STAGES]
    uint64_t* bar_q2 = bar_tempty + ACC_STAGES;        // [1] hybrid: term q2 has landed in shared memory
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_q2 + 1);
    float* s_aux_all = reinterpret_cast<float*>(s_tail + 256);   // [4 * EW epilogue warps][NV]: per-row constants of the warp's current column group

    const uint32_t q0 = blockIdx.x * BM;
    // contiguous: split y owns rows [y * rows_per_split, ...); strided: tiles y, y + n_splits, y + 2 n_splits, ... of the database
    const uint32_t tile_step = p.strided ? p.n_splits : 1u;
    const uint64_t r_begin = p.strided ? static_cast<uint64_t>(blockIdx.y) * BN : static_cast<uint64_t>(blockIdx.y) * p.rows_per_split;
    const uint64_t r_end = p.strided ? static_cast<uint64_t>(p.n_pad) : min(static_cast<uint64_t>(p.n_pad), r_begin + p.rows_per_split);
    const uint32_t n_tiles = (r_begin < r_end) ? static_cast<uint32_t>((r_end - r_begin + static_cast<uint64_t>(BN) * tile_step - 1) / (static_cast<uint64_t>(BN) * tile_step)) : 0;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.n_stages; s++) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }
        mbar_init(bar_q, TS ? EPI_T : 1);
        mbar_init(bar_q2, 1);
        for (uint32_t a = 0; a < NACC; a++) { mbar_init(bar_tfull + a, 1); mbar_init(bar_tempty + a, EPI_T); }
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
            if (!TS && !stream_q) {
                mbar_expect_tx(bar_q, p.a_pieces * p.nslab * SLAB_TILE);
                for (uint32_t a = 0; a < p.a_pieces; a++)
                    for (uint32_t s = 0; s < p.nslab; s++)
                        tma_load_2d(smem_u32(s_q + (static_cast<size_t>(a) * p.nslab + s) * SLAB_TILE), &tm_q, bar_q, s * SLAB_ELEMS,
                                    a * p.nq_pad + q0);
            }
            if (hyb || lo_s) {
                const uint32_t piece = hyb ? 2u : 1u;
                mbar_expect_tx(bar_q2, p.nslab * SLAB_TILE);
                for (uint32_t s = 0; s < p.nslab; s++)
                    tma_load_2d(smem_u32(s_q + static_cast<size_t>(s) * SLAB_TILE), &tm_q, bar_q2, s * SLAB_ELEMS, piece * p.nq_pad + q0);
            }
            uint32_t stage = 0, ph = 0;          // ring position and its phase bit, advanced incrementally (no division per slab)
            long long w_prod = 0;
            for (uint32_t t = 0; t < n_tiles; t++) {
                const uint32_t row0 = static_cast<uint32_t>(r_begin) + t * tile_step * BN;
                for (uint32_t s = 0; s < p.nslab; s++, ph ^= (++stage == p.n_stages) ? 1u : 0u, stage = (stage == p.n_stages) ? 0u : stage) {
                    mbar_wait_timed(bar_empty + stage, ph ^ 1u, w_prod);
                    mbar_expect_tx(bar_full + stage, stage_slabs * SLAB_TILE);
                    for (uint32_t a = 0; a < q_slabs; a++)
                        tma_load_2d(smem_u32(s_x + (static_cast<size_t>(stage) * stage_slabs + a) * SLAB_TILE), &tm_q, bar_full + stage, s * SLAB_ELEMS,
                                    a * p.nq_pad + q0);
                    for (int b = 0; b < NB; b++)
                        tma_load_2d(smem_u32(s_x + (static_cast<size_t>(stage) * stage_slabs + q_slabs + b) * SLAB_TILE), &tm_x, bar_full + stage, s * SLAB_ELEMS,
                                    b * p.n_pad + row0);
                }
            }
            if (TC_COUNTERS && p.dbg_cycles && blockIdx.x == 0 && blockIdx.y == 0) p.dbg_cycles[1] = w_prod;
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        // The whole warp walks the pipeline converged (so every address / descriptor is warp-uniform and lives in
        // uniform registers); only the tcgen05 instructions themselves are issued by one elected lane.
        constexpr uint32_t idesc = make_idesc(KIND);
        constexpr uint32_t SLAB_DESC = SLAB_TILE >> 4;    // descriptor start-address units (16 B) per slab
        if (!stream_q) mbar_wait(bar_q, 0);
        if (hyb || lo_s) mbar_wait(bar_q2, 0);
        tc_fence_after();
        const uint32_t q_desc0 = make_smem_desc(smem_u32(s_q));   // low descriptor words; the constant high word is added by umma()
        const uint32_t x_desc0 = make_smem_desc(smem_u32(s_x));
        uint32_t stage = 0, ph = 0, acc = 0, aph = 0;   // ring / accumulator positions and phase bits, advanced incrementally
        long long w_full = 0, w_tempty = 0;
        const long long t_start = tc_clock();
        for (uint32_t t = 0; t < n_tiles; t++, aph ^= (++acc == NACC) ? 1u : 0u, acc = (acc == NACC) ? 0u : acc) {
            mbar_wait_timed(bar_tempty + acc, aph ^ 1u, w_tempty);
            tc_fence_after();
            const uint32_t tmem_c = tmem_base + ACC_COL0 + acc * BN;
            for (uint32_t s = 0; s < p.nslab; s++, ph ^= (++stage == p.n_stages) ? 1u : 0u, stage = (stage == p.n_stages) ? 0u : stage) {
                mbar_wait_timed(bar_full + stage, ph, w_full);
                tc_fence_after();
                const uint32_t xd = x_desc0 + (stage * stage_slabs + q_slabs) * SLAB_DESC;
                const uint32_t qd = stream_q ? x_desc0 + stage * stage_slabs * SLAB_DESC : q_desc0 + s * SLAB_DESC;
                const uint32_t q_piece = stream_q ? SLAB_DESC : p.nslab * SLAB_DESC;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < KSTEPS; k++) {
                        const uint32_t first = (s | static_cast<uint32_t>(k)) != 0 ? 1u : 0u;   // 0 only for the tile's first MMA
                        if (TS && SPLIT3) {
                            // queries in TMEM: hi at columns [0, PIECE_COLS), lo behind it; 8 columns (8 tf32 / 16 fp16) per K step
                            const uint32_t a_hi = tmem_base + s * 32 + k * 8, a_lo = a_hi + PIECE_COLS;
                            umma_ts<KIND>(tmem_c, a_hi, xd + 2 * k, idesc, first);
                            if (lo_s) umma<KIND>(tmem_c, qd + 2 * k, xd + 2 * k, idesc, 1u);     // Qlo from shared memory
                            else umma_ts<KIND>(tmem_c, a_lo, xd + 2 * k, idesc, 1u);
                            umma_ts<KIND>(tmem_c, a_hi, xd + SLAB_DESC + 2 * k, idesc, 1u);
                        } else if (TS && KIND == KIND_I8) {
                            // int8 codes, four per column: 8 columns (32 codes) per K step; one exact s32 term
                            umma_ts<KIND>(tmem_c, tmem_base + s * 32 + k * 8, xd + 2 * k, idesc, first);
                        } else if (TS) {
                            // bf16 query terms q0, q1, q2 in TMEM at columns [0,64), [64,128), [128,192); 8 columns (16 bf16) per K step
                            const uint32_t a0 = tmem_base + s * 32 + k * 8;
                            umma_ts<KIND>(tmem_c, a0, xd + 2 * k, idesc, first);
                            if (p.a_pieces > 1) umma_ts<KIND>(tmem_c, a0 + PIECE_COLS, xd + 2 * k, idesc, 1u);
                            if (p.a_pieces > 2) {
                                if (hyb) umma<KIND>(tmem_c, qd + 2 * k, xd + 2 * k, idesc, 1u);
                                else umma_ts<KIND>(tmem_c, a0 + 2 * PIECE_COLS, xd + 2 * k, idesc, 1u);
                            }
                        } else if (SPLIT3) {
                            // s = Qhi.Xhi + Qlo.Xhi + Qhi.Xlo   (the lo.lo term is below 2^-22 relative)
                            umma<KIND>(tmem_c, qd + 2 * k, xd + 2 * k, idesc, first);
                            umma<KIND>(tmem_c, qd + q_piece + 2 * k, xd + 2 * k, idesc, 1u);
                            umma<KIND>(tmem_c, qd + 2 * k, xd + SLAB_DESC + 2 * k, idesc, 1u);
                        } else {
                            umma<KIND>(tmem_c, qd + 2 * k, xd + 2 * k, idesc, first);
                            if (p.a_pieces > 1) umma<KIND>(tmem_c, qd + q_piece + 2 * k, xd + 2 * k, idesc, 1u);
                            if (p.a_pieces > 2) umma<KIND>(tmem_c, qd + 2 * q_piece + 2 * k, xd + 2 * k, idesc, 1u);
                        }
                    }
                    umma_commit(bar_empty + stage);                       // slab consumed once the MMAs above retire
                    if (s + 1 == p.nslab) umma_commit(bar_tfull + acc);   // accumulator ready for the epilogue
                }
                __syncwarp();
            }
        }
        if (TC_COUNTERS && p.dbg_cycles && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) {
            p.dbg_cycles[0] = tc_clock() - t_start;
            p.dbg_cycles[2] = w_full;
            p.dbg_cycles[3] = w_tempty;
            p.dbg_cycles[6] = n_tiles;
        }
        __syncwarp();
    } else {
        // ===================================================================== epilogue (4 warps, thread = query row)
        const uint32_t quarter = warp & 3u;                 // TMEM lane quarter this warp may access
        const uint32_t row_in_tile = quarter * 32 + lane;   // query row inside the tile
        const uint32_t half = (warp - 2) >> 2;              // which NV-column group of every tile this warp scans (0 .. EW - 1)
        TopList<KP> top;
        top.init();
        float scratch[NV];
        long long w_tfull = 0, w_slow = 0;
        // Shared threshold: the k'-th best value any CTA of this query has seen so far (monotone, atomicMin on the
        // order-preserving integer image).  A value that does not beat it cannot be in the merged top-k', whichever
        // split holds it, so every split prunes with the tightest bound known anywhere.  Stale reads are only looser.
        if (TS) {
            // Stage this CTA's 128 queries into TMEM once: query pieces are dealt to the two warps of each lane quarter
            // (even pieces to half 0, odd to half 1); thread = query row = TMEM lane, 32 columns (128 B of the row) per
            // tcgen05.st.  16-bit operands sit two per column, low half = even k, exactly as in memory.
            const uint32_t row_words = p.kp * ELEM / 4;
            const uint32_t tmem_pieces = hyb ? 2u : (lo_s ? 1u : p.a_pieces);
            const uint32_t cpp = row_words / 32;                       // 32-column chunks per piece
            for (uint32_t item = half; item < tmem_pieces * cpp; item += EW) {   // (piece, chunk) items dealt round-robin to the quarter's warps
                const uint32_t pc = item / cpp, c = (item - pc * cpp) * 32;
                const uint32_t* src = reinterpret_cast<const uint32_t*>(p.q_op) + (static_cast<size_t>(pc) * p.nq_pad + q0 + row_in_tile) * row_words;
                const uint32_t tq = tmem_base + ((quarter * 32u) << 16) + pc * PIECE_COLS;
                uint32_t w[32];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint4 x = __ldg(reinterpret_cast<const uint4*>(src + c) + j);
                    w[4 * j] = x.x; w[4 * j + 1] = x.y; w[4 * j + 2] = x.z; w[4 * j + 3] = x.w;
                }
                tmem_st32(tq + c, w);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_q);
        }
        uint32_t* gtau_ptr = p.gtau + q0 + row_in_tile;
        const bool share_tau = !DENSE && p.wide_k == 0;     // wide-k lists prune with their own thresholds only (see Params::wide_k)
        uint32_t g_next = share_tau ? *reinterpret_cast<volatile uint32_t*>(gtau_ptr) : 0xFFFFFFFFu;
        // Per-column constants of this warp's 64-column half: lane l fetches columns l and l + 32 (coalesced, one tile
        // ahead) and parks them in a warp-private shared-memory row; the value loop reads them back as 128-bit broadcast
        // loads (16 per tile instead of 64 shuffles).  No CTA-wide barrier: only __syncwarp.
        // UNIT (flat f32 index, cosine, 3xFP16): the database operand holds the rows divided by their index norm (uniform scale 2^13;
        // pad rows and zero-norm rows carry a NaN, which min / compare ignore exactly as the NaN row constant used to) and the query
        // operand is negated, so the accumulator IS the selection value up to the query's power-of-two scale: no per-column constant,
        // no multiply, no staging of constants on the per-tile chain.
        constexpr bool UNIT = (KIND == KIND_F16X3) && (MET != MET_L2) && !DENSE;
        float* s_aux = s_aux_all + (warp - 2) * NV;
        const float* aux_half = p.aux + r_begin + half * NV + lane;
        const size_t aux_step = static_cast<size_t>(tile_step) * BN;
        float aux_lo_next = 0.f, aux_hi_next = 0.f;
        if (!UNIT && n_tiles > 0) { aux_lo_next = __ldg(aux_half); if (NV > 32) aux_hi_next = __ldg(aux_half + 32); }
        // 3xFP16: operands were scaled by powers of two -- queries per row (cq undoes it), the database per index for L2
        // (v = |x|^2 - 2 s / (sq S): folded into cq as well) and per row for the cosine forms that keep a row constant
        // (the DENSE centroid values: v = s * (-1 / (|c| sc)) / sq); the flat cosine operand needs nothing (UNIT, above).
        constexpr bool RX = false;      // (L2 operands carry one uniform scale since the closing rework: nothing per column to undo)
        float* s_rx = s_aux_all + (4 * EW + (warp - 2)) * NV;           // second bank of warp-private rows
        const float* rx_half = RX ? p.aux2 + r_begin + half * NV + lane : nullptr;
        float rx_lo_next = 1.f, rx_hi_next = 1.f;
        if (RX && n_tiles > 0) { rx_lo_next = __ldg(rx_half); if (NV > 32) rx_hi_next = __ldg(rx_half + 32); }
        float cq = 1.0f;
        if (KIND == KIND_F16X3) cq = __ldg(p.q_inv_scale + q0 + row_in_tile) * ((MET == MET_L2) ? -2.0f * p.db_inv_scale : (UNIT ? 1.0f / 8192.0f : 1.0f));
        // 3xFP16 cosine: cq is a positive power of two, so w = v / cq orders exactly like v and converts back without rounding.  The
        // candidate list and the threshold test run on w = s * aux -- one multiply per value instead of two on the per-tile chain --
        // and values cross into the common domain (shared threshold, emitted keys) through exact multiplications by cq / its inverse.
        constexpr bool WDOM = (KIND == KIND_F16X3) && (MET != MET_L2) && !DENSE;
        const float inv_cq = WDOM ? 1.0f / cq : 1.0f;
        uint32_t acc = 0, aph = 0;               // accumulator stage and its phase bit, advanced incrementally (no division per tile)
        for (uint32_t t = 0; t < n_tiles; t++, aph ^= (++acc == NACC) ? 1u : 0u, acc = (acc == NACC) ? 0u : acc) {
            const uint32_t row0 = static_cast<uint32_t>(r_begin) + t * tile_step * BN;
            const uint32_t g_bits = g_next;
            const float aux_lo = aux_lo_next, aux_hi = aux_hi_next;
            const float rx_lo = rx_lo_next, rx_hi = rx_hi_next;
            if (t + 1 < n_tiles) {   // prefetch for the next tile: latency hidden behind this tile's work
                if (!UNIT) {
                    aux_lo_next = __ldg(aux_half + static_cast<size_t>(t + 1) * aux_step);
                    if (NV > 32) aux_hi_next = __ldg(aux_half + static_cast<size_t>(t + 1) * aux_step + 32);
                }
                if (RX) {
                    rx_lo_next = __ldg(rx_half + static_cast<size_t>(t + 1) * aux_step);
                    if (NV > 32) rx_hi_next = __ldg(rx_half + static_cast<size_t>(t + 1) * aux_step + 32);
                }
                if (share_tau) g_next = *reinterpret_cast<volatile uint32_t*>(gtau_ptr);
            }
            if (!UNIT) {
                __syncwarp();                      // every lane is done with the previous tile's constants
                s_aux[lane] = aux_lo;
                if (NV > 32) s_aux[lane + 32] = aux_hi;
                if (RX) { s_rx[lane] = rx_lo; if (NV > 32) s_rx[lane + 32] = rx_hi; }
                __syncwarp();
            }
            mbar_wait_timed(bar_tfull + acc, aph, w_tfull);
            tc_fence_after();
            const float g_tau = WDOM ? ordered_to_f32(g_bits) * inv_cq : ordered_to_f32(g_bits);       // NaN (all-ones init) is ignored by fminf
            float tau = fminf(top.tau(), g_tau);
            const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + ACC_COL0 + acc * BN;
            {
                const int c = static_cast<int>(half);
                uint32_t r[NV];
                if constexpr (NV == 64) tmem_ld64_sync(taddr + c * NV, r);
                else tmem_ld32_sync(taddr + c * NV, r);
                // the tile's values are in registers: hand the accumulator stage back before the select work, so the
                // issuer never waits on this warp's candidate inserts
                tc_fence_before();
                mbar_arrive(bar_tempty + acc);
                float v[NV];
                float gm[NG];  // minima of the groups of 8 columns
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    float mg = INFINITY;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int col = g * 8 + j;   // compile-time after unrolling
                        const float cst = UNIT ? 0.0f : s_aux[col];
                        const float sdot = (KIND == KIND_I8) ? __int2float_rn(static_cast<int32_t>(r[col])) : __uint_as_float(r[col]);   // s32 dots are exact in f32 (< 2^24)
                        if (UNIT) v[col] = sdot;
                        else if (KIND == KIND_F16X3) v[col] = (MET == MET_L2) ? fmaf(sdot, cq, cst) : (WDOM ? sdot * cst : (sdot * cst) * cq);
                        else v[col] = (MET == MET_L2) ? fmaf(sdot, -2.0f, cst) : sdot * cst;
                        mg = fminf(mg, v[g * 8 + j]);
                    }
                    gm[g] = mg;
                }
                if constexpr (DENSE) {
                    if (static_cast<uint64_t>(q0) + row_in_tile < p.nq) {
                        if (p.dense_blocked) {
                            float4* dst = reinterpret_cast<float4*>(p.dense) + dense_piece_index(static_cast<uint64_t>(q0) + row_in_tile, (row0 + c * NV) >> 2, p.dense_ld >> 7);
#pragma unroll
                            for (int j = 0; j < NV / 4; j++) dst[j * 128] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        } else {
                            float4* dst = reinterpret_cast<float4*>(p.dense + (static_cast<uint64_t>(q0) + row_in_tile) * p.dense_ld + row0 + c * NV);
#pragma unroll
                            for (int j = 0; j < NV / 4; j++) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        }
                        if (p.dense_gm != nullptr) {   // group minima (pad columns are +inf / NaN and never lower a minimum)
                            float4* gdst = reinterpret_cast<float4*>(p.dense_gm + (static_cast<uint64_t>(q0) + row_in_tile) * (p.dense_ld >> 3) + ((row0 + c * NV) >> 3));
#pragma unroll
                            for (int j = 0; j < NG / 4; j++) gdst[j] = make_float4(gm[4 * j], gm[4 * j + 1], gm[4 * j + 2], gm[4 * j + 3]);
                        }
                    }
                }
                float m = INFINITY;
                if (!DENSE) {
#pragma unroll
                    for (int g = 0; g < NG; g++) m = fminf(m, gm[g]);
                }
                if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && t == 0) {
#pragma unroll
                    for (int j = 0; j < NV; j++) p.dbg[row_in_tile * BN + c * NV + j] = WDOM ? v[j] * cq : v[j];
                }
                if (m < tau) {
                    const long long ts0 = tc_clock();
                    select_from_tile<KP, NV>(top, tau, v, gm, m, row0 + c * NV, scratch);
                    w_slow += tc_clock() - ts0;
                }
            }
            if (share_tau && (top.tau() < g_tau || (g_tau != g_tau && top.tau() < INFINITY))) atomicMin(gtau_ptr, f32_to_ordered(WDOM ? top.tau() * cq : top.tau()));
        }
        if (!DENSE && !share_tau && top.tau() < INFINITY) atomicMin(gtau_ptr, f32_to_ordered(WDOM ? top.tau() * cq : top.tau()));   // wide-k: this list's final threshold
        if (TC_COUNTERS && p.dbg_cycles && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64) { p.dbg_cycles[4] = w_tfull; p.dbg_cycles[5] = w_slow; }
        const uint64_t q = static_cast<uint64_t>(q0) + row_in_tile;
        if (!DENSE && q < p.nq) {
            uint64_t* out = p.part_keys + (q * (EW * p.n_splits) + EW * blockIdx.y + half) * KP;
#pragma unroll
            for (int j = 0; j < KP; j++) out[j] = (top.i[j] == IDX_INVALID) ? KEY_SENTINEL : make_key(WDOM ? top.v[j] * cq : top.v[j], top.i[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ----------------------------------------------------------------------------------------------- IVF centroid ranking
// Second stage of the tensor-core centroid ranking (src/cpu/ivf.rs:349-365 get_centroids_dist + select_probed_clusters'
// sort): one warp per query turns its row of approximate values (DENSE mode of the kernel above) into the `pitch`
// nearest (distance, cell) keys in ascending order, exact in the reference's arithmetic:
//   1. radix select (4 passes of 8 bits over the order-preserving integer image) finds the pitch-th smallest value,
//   2. every cell at or below it becomes a candidate (ties included; more than cmax candidates -> nothing is emitted),
//   3. the candidates' distances are recomputed with refdist.cuh and sorted by (distance, cell),
//   4. only ranks that provably precede every non-candidate are emitted: exact distance < threshold - error bound.
// Ranks that cannot be certified are written as sentinels; probe_walk_kernel flags a query that runs into one and the
// caller repeats the batch on the exact CUDA-core ranking.
struct CoarseSelectParams {
    const float* dense;         // [nq][dense_ld] approximate selection values
    uint32_t dense_ld;
    uint32_t blocked;           // dense is stored in the blocked layout of dense_piece_index()
    const float* gmin;          // coarse_select_gm_kernel: [nq][dense_ld / 8] minima of the aligned groups of 8 values
    uint32_t glimit;            // ... its radix select stops once at most this many groups lie at or below the threshold
    uint64_t nq;
    uint32_t nlist, pitch, cmax;   // cmax: candidate capacity, a power of two >= pitch
    uint32_t staged_words;         // round_up(nlist, 4) when the row of values is kept in shared memory, 0 otherwise
    const float* queries;       // f32 routing queries [nq][q_ld]
    uint32_t q_ld;
    const float* centroids;     // [nlist][c_ld]
    uint32_t c_ld;
    const float* centroid_norms;
    uint32_t dim;
    float eps, cnorm_max;
    uint64_t* ranked;           // [nq][pitch]
    uint32_t do_walk;           // 1: apply the probe-expansion rule to the certified prefix here and write the probe lists instead of `ranked`
    CoarseWalk walk;
};

__host__ __device__ inline size_t coarse_select_warp_bytes(uint32_t cmax, uint32_t staged_words) {
    return static_cast<size_t>(cmax) * 8 + 256 * 4 + static_cast<size_t>(cmax) * 4 + static_cast<size_t>(staged_words) * 4;   // keys | histogram | candidate cells | staged row
}
// coarse_select_gm_kernel: keys | candidate cells | area shared by (histogram + group minima) and (8 staged centroid rows + query row).
// Staged rows sit 8 (mod 32) words apart: the 8-byte reads of a half warp (4 rows x 4 lanes) fall into distinct banks.
__host__ __device__ inline uint32_t coarse_stage_row_words(uint32_t c_ld) { return ((c_ld + 31u) & ~31u) + 8u; }
__host__ __device__ inline size_t coarse_stage_bytes(uint32_t c_ld, uint32_t dim) { return 8ull * coarse_stage_row_words(c_ld) * 4 + ((dim * 4ull + 15) & ~15ull); }
__host__ __device__ inline size_t coarse_gm_warp_bytes(uint32_t cmax, uint32_t ngrp, uint32_t c_ld, uint32_t dim) {
    const size_t sel = 256 * 4 + static_cast<size_t>(ngrp) * 4, dist = coarse_stage_bytes(c_ld, dim);
    return static_cast<size_t>(cmax) * 12 + ((sel > dist ? sel : dist) + 15) / 16 * 16;
}

// Radix select over order-preserving integer images, one warp: a threshold thr such that the valid values at or below it
// number at least `need` and -- as soon as a digit boundary allows -- at most `limit`.  Most significant digit first, starting
// at the first bit in which the values differ (the common high bits -- sign, exponent -- would put every value into one
// histogram bin); it stops as soon as the values at or below the chosen bin fit `limit` (the bin's upper edge is the
// threshold) or no bits are left (ties: the count may then exceed `limit`, the caller checks).
// fetch(i4, u): images 4 i4 .. 4 i4 + 3; values at positions >= nvalid are ignored.
template <class Fetch>
__device__ __forceinline__ uint32_t coarse_radix_threshold(Fetch fetch, uint32_t n4, uint32_t nvalid, uint32_t need, uint32_t limit, uint32_t umin,
                                                           uint32_t umax, uint32_t* hist, uint32_t lane) {
    uint32_t rem = 32u - __clz(umin ^ umax);          // unresolved low bits (0: all values equal)
    uint32_t prefix = rem >= 32u ? 0u : (umin >> rem) << rem;
    uint32_t below = 0;
    uint32_t thr = prefix;                             // rem == 0
    while (rem > 0) {
        const uint32_t w = min(8u, rem), sh = rem - w, dmask = (1u << w) - 1u;
        const uint32_t hmask = rem >= 32u ? 0u : ~0u << rem;
        for (uint32_t b = lane; b < 256; b += 32) hist[b] = 0;
        __syncwarp();
        for (uint32_t i4 = lane; i4 < n4; i4 += 32) {
            uint32_t u[4];
            fetch(i4, u);
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (4 * i4 + e < nvalid && (u[e] & hmask) == prefix) atomicAdd(hist + ((u[e] >> sh) & dmask), 1u);
        }
        __syncwarp();
        uint32_t h[8], local = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { h[j] = hist[8 * lane + j]; local += h[j]; }
        uint32_t incl = local;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
            if (lane >= static_cast<uint32_t>(off)) incl += t;
        }
        const uint32_t excl = incl - local;
        const bool mine = excl < need && need <= incl;
        const int owner = __ffs(__ballot_sync(0xFFFFFFFFu, mine)) - 1;
        uint32_t bin = 0, before = 0, upto = 0;
        if (mine) {
            uint32_t c = excl;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (need > c && need <= c + h[j]) { bin = 8 * lane + j; before = c; upto = c + h[j]; }
                c += h[j];
            }
        }
        bin = __shfl_sync(0xFFFFFFFFu, bin, owner);
        before = __shfl_sync(0xFFFFFFFFu, before, owner);
        upto = __shfl_sync(0xFFFFFFFFu, upto, owner);
        __syncwarp();
        if (below + upto <= limit || sh == 0) {      // few enough values at or below this bin (or nothing left to refine)
            thr = prefix | (bin << sh) | ((sh ? (1u << sh) : 1u) - 1u);
            break;
        }
        below += before;
        need -= before;
        prefix |= bin << sh;
        rem = sh;
    }
    return thr;
}

// Threshold and candidates from the full row of values (read from L2 by every pass): every cell whose value is <= thr.
// Returns the candidate count (cells beyond cmax are counted, not stored).
__device__ __forceinline__ uint32_t coarse_candidates_full_row(const CoarseSelectParams& p, uint64_t q, uint32_t* hist, uint32_t* cand, uint4* staged,
                                                               uint32_t lane, uint32_t& thr) {
    const float4* src4 = reinterpret_cast<const float4*>(p.dense + q * p.dense_ld);
    const float4* blk4 = reinterpret_cast<const float4*>(p.dense);
    const uint32_t col_tiles = p.dense_ld >> 7;
    auto piece = [&](uint32_t i4) { return p.blocked ? __ldg(blk4 + dense_piece_index(q, i4, col_tiles)) : __ldg(src4 + i4); };
    const uint32_t n4 = (p.nlist + 3) >> 2;
    const bool use_smem = staged != nullptr;
    auto ord = [&](float4 x, uint32_t i4, uint32_t u[4]) {   // ordered images; columns past nlist never qualify
        u[0] = f32_to_ordered(x.x); u[1] = f32_to_ordered(x.y); u[2] = f32_to_ordered(x.z); u[3] = f32_to_ordered(x.w);
#pragma unroll
        for (int e = 0; e < 4; e++) if (4 * i4 + e >= p.nlist) u[e] = 0xFFFFFFFFu;
    };
    auto fetch = [&](uint32_t i4, uint32_t u[4]) {
        if (use_smem) { const uint4 w = staged[i4]; u[0] = w.x; u[1] = w.y; u[2] = w.z; u[3] = w.w; }
        else ord(piece(i4), i4, u);
    };
    thr = 0xFFFFFFFFu;
    if (p.pitch < p.nlist) {
        uint32_t umin = 0xFFFFFFFFu, umax = 0u;
        for (uint32_t i4 = lane; i4 < n4; i4 += 32) {
            uint32_t u[4];
            ord(piece(i4), i4, u);
            if (use_smem) staged[i4] = make_uint4(u[0], u[1], u[2], u[3]);
#pragma unroll
            for (int e = 0; e < 4; e++) if (4 * i4 + e < p.nlist) { umin = min(umin, u[e]); umax = max(umax, u[e]); }
        }
        __syncwarp();
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            umin = min(umin, __shfl_xor_sync(0xFFFFFFFFu, umin, off));
            umax = max(umax, __shfl_xor_sync(0xFFFFFFFFu, umax, off));
        }
        thr = coarse_radix_threshold(fetch, n4, p.nlist, p.pitch, p.cmax, umin, umax, hist, lane);
    } else if (use_smem) {
        for (uint32_t i4 = lane; i4 < n4; i4 += 32) {
            uint32_t u[4];
            ord(piece(i4), i4, u);
            staged[i4] = make_uint4(u[0], u[1], u[2], u[3]);
        }
        __syncwarp();
    }
    uint32_t count = 0;
    for (uint32_t i0 = 0; i0 < n4; i0 += 32) {
        const uint32_t i4 = i0 + lane;
        uint32_t u[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        if (i4 < n4) fetch(i4, u);
        uint32_t mine = 0;
#pragma unroll
        for (int e = 0; e < 4; e++) mine += (4 * i4 + e < p.nlist && u[e] <= thr) ? 1u : 0u;
        uint32_t incl = mine;   // inclusive prefix of the per-lane hit counts
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
            if (lane >= static_cast<uint32_t>(off)) incl += t;
        }
        uint32_t pos = count + incl - mine;
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (4 * i4 + e < p.nlist && u[e] <= thr) { if (pos < p.cmax) cand[pos] = 4 * i4 + e; pos++; }
        count += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    __syncwarp();
    return count;
}

// Steps 3 and 4 of the centroid select: exact distances of the candidates (reference arithmetic), (distance, cell) order,
// certified prefix.  thr: every cell that is NOT a candidate has an approximate value above it.
// stage != nullptr: a warp-private shared-memory area of coarse_stage_bytes(): the candidates' centroid rows are brought in eight
// at a time with coalesced 16-byte cp.async copies (all of a batch's loads in flight together), and four lanes share one
// candidate: lane t of a group owns SIMD lanes 2t and 2t + 1 of the reference's 8-lane accumulation (src/utils/dist.rs:306-330,
// 587-609), walks the chunks in order with the same rounding steps, and the group folds with the reference's horizontal-add
// tree by shuffles -- the bits of one thread walking the row (refdist.cuh), a ninth of the dependent L2 round trips.
template <int MET>
__device__ __forceinline__ void coarse_emit_ranked(const CoarseSelectParams& p, uint64_t q, uint32_t thr, uint32_t count, uint64_t* keys,
                                                   const uint32_t* cand, uint32_t lane, uint8_t* stage = nullptr) {
    uint64_t* out = p.ranked + q * p.pitch;
    if (count > p.cmax) {   // a huge tie class: leave the query to the exact ranking
        if (p.do_walk) {
            if (lane == 0) { atomicExch(p.walk.overflow, 1u); p.walk.n_probes[q] = 0; }
            return;
        }
        for (uint32_t j = lane; j < p.pitch; j += 32) out[j] = KEY_SENTINEL;
        return;
    }
    const uint8_t* qrow = reinterpret_cast<const uint8_t*>(p.queries + q * p.q_ld);
    if (stage != nullptr) {
        const uint32_t rs = coarse_stage_row_words(p.c_ld) * 4;
        float* qbuf = reinterpret_cast<float*>(stage + 8 * rs);
        __syncwarp();                                     // the selection is done with this area (histogram, group minima)
        for (uint32_t e = lane; e < p.dim; e += 32) qbuf[e] = reinterpret_cast<const float*>(qrow)[e];
        __syncwarp();
        qrow = reinterpret_cast<const uint8_t*>(qbuf);
    }
    float qn = 1.0f;
    if (MET == MET_COS) {
        if (lane == 0) qn = seq_norm<4>(qrow, p.dim);
        qn = __shfl_sync(0xFFFFFFFFu, qn, 0);
    }
    if (stage != nullptr) {
        const uint32_t rs = coarse_stage_row_words(p.c_ld) * 4, row_bytes = p.c_ld * 4;
        const uint32_t t = lane & 3u, r = lane >> 2;
        const uint32_t chunks = p.dim >> 3;
        for (uint32_t base = 0; base < count; base += 8) {
            const uint32_t nb = min(8u, count - base);
            for (uint32_t i = 0; i < nb; i++) {
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.centroids + static_cast<uint64_t>(cand[base + i]) * p.c_ld);
                for (uint32_t b = lane * 16; b < row_bytes; b += 512) cp_async16(stage + i * rs + b, src + b);
            }
            cp_async_commit();
            cp_async_wait<0>();
            __syncwarp();
            const float* x = reinterpret_cast<const float*>(stage + r * rs);
            const float* y = reinterpret_cast<const float*>(qrow);
            float a0 = 0.0f, a1 = 0.0f;
            if (r < nb) {
#pragma unroll 4
                for (uint32_t c = 0; c < chunks; c++) {
                    const float2 xv = *reinterpret_cast<const float2*>(x + 8 * c + 2 * t), yv = *reinterpret_cast<const float2*>(y + 8 * c + 2 * t);
                    if (MET == MET_L2) {
                        const float d0 = __fsub_rn(xv.x, yv.x), d1 = __fsub_rn(xv.y, yv.y);
                        a0 = __fadd_rn(a0, __fmul_rn(d0, d0));
                        a1 = __fadd_rn(a1, __fmul_rn(d1, d1));
                    } else {
                        a0 = __fadd_rn(a0, __fmul_rn(xv.x, yv.x));
                        a1 = __fadd_rn(a1, __fmul_rn(xv.y, yv.y));
                    }
                }
            }
            // wide::f32x8::reduce_add: s_j = a_j + a_(j+4); (s0 + s2) + (s1 + s3)
            const float s0 = __fadd_rn(a0, __shfl_down_sync(0xFFFFFFFFu, a0, 2, 4)), s1 = __fadd_rn(a1, __shfl_down_sync(0xFFFFFFFFu, a1, 2, 4));   // lane 0: s0, s1; lane 1: s2, s3
            const float u0 = __fadd_rn(s0, __shfl_down_sync(0xFFFFFFFFu, s0, 1, 4)), u1 = __fadd_rn(s1, __shfl_down_sync(0xFFFFFFFFu, s1, 1, 4));
            if (t == 0 && r < nb) {
                float sum = __fadd_rn(u0, u1);
                for (uint32_t e = chunks * 8; e < p.dim; e++) {   // scalar tail: `sum += d * d` (not fused)
                    if (MET == MET_L2) {
                        const float d = __fsub_rn(x[e], y[e]);
                        sum = __fadd_rn(sum, __fmul_rn(d, d));
                    } else {
                        sum = __fadd_rn(sum, __fmul_rn(x[e], y[e]));
                    }
                }
                const uint32_t c = cand[base + r];
                const float cn = (MET == MET_COS) ? p.centroid_norms[c] : 1.0f;
                keys[base + r] = make_key(finish_fp<MET>(sum, qn, cn), c);
            }
            __syncwarp();                                 // the rows are consumed before the next batch lands on them
        }
        for (uint32_t j = count + lane; j < p.cmax; j += 32) keys[j] = KEY_SENTINEL;
    } else
    for (uint32_t j = lane; j < p.cmax; j += 32) {
        uint64_t key = KEY_SENTINEL;
        if (j < count) {
            const uint32_t c = cand[j];
            float raw[1];
            accumulate_fp<4, 4, MET == MET_L2, 1, true>(reinterpret_cast<const uint8_t*>(p.centroids + static_cast<uint64_t>(c) * p.c_ld), qrow, p.q_ld * 4, p.dim, raw);
            const float cn = (MET == MET_COS) ? p.centroid_norms[c] : 1.0f;
            key = make_key(finish_fp<MET>(raw[0], qn, cn), c);
        }
        keys[j] = key;
    }
    __syncwarp();
    if (p.cmax == 128u) {          // the usual capacity (prefixes of up to 127 ranks): sort in registers, shuffles instead of shared-memory passes
        uint64_t kr[4];
#pragma unroll
        for (int j = 0; j < 4; j++) kr[j] = keys[lane + 32 * j];
        warp_bitonic_sort_regs<4>(kr, lane);
#pragma unroll
        for (int j = 0; j < 4; j++) keys[lane + 32 * j] = kr[j];
        __syncwarp();
    } else {
        bitonic_sort_keys<false>(keys, p.cmax, lane, 32);
    }
    float bound = INFINITY;
    if (p.pitch < p.nlist) {
        double qn2 = 0.0;   // f64: the value -> distance map must not add rounding of its own
        for (uint32_t e = lane; e < p.dim; e += 32) {
            const double x = reinterpret_cast<const float*>(qrow)[e];
            qn2 += x * x;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) qn2 += __shfl_xor_sync(0xFFFFFFFFu, qn2, off);
        const double tv = ordered_to_f32(thr);   // every non-candidate's approximate value exceeds this
        double b;
        if (MET == MET_L2) {
            const double s = sqrt(qn2) + p.cnorm_max;            // value = dist - |q|^2, error <= eps (|q| + |c|max)^2
            b = tv + qn2 - p.eps * s * s;
        } else if (MET == MET_COS) {
            const double qq = qn;                                 // value = (dist - 1) |q| with the reference's |q|, error <= eps
            b = qq > 0.0 ? tv / qq + 1.0 - p.eps : -INFINITY;
        } else {
            b = tv + 1.0 - p.eps * fmax(1.0, sqrt(qn2) * p.cnorm_max);   // pre-normalised: value = dist - 1
        }
        bound = __double2float_rd(b);
    }
    if (!p.do_walk) {
        for (uint32_t j = lane; j < p.pitch; j += 32) {
            const uint64_t key = keys[j];
            out[j] = (key_idx(key) != IDX_INVALID && key_dist(key) < bound) ? key : KEY_SENTINEL;
        }
        return;
    }
    // Probe expansion on the certified prefix, 32 ranks per step: cells are taken in rank order until nprobe are chosen AND at least
    // k vectors are reachable (empty cells count toward nprobe only); an uncertified rank ends the prefix (probe_walk_kernel's rule).
    uint32_t chosen = 0;
    unsigned long long reach = 0, local = 0;
    bool done = false;
    for (uint32_t base = 0; base < p.pitch; base += 32) {
        const uint32_t i = base + lane;
        uint32_t c = IDX_INVALID;
        if (i < p.pitch) {
            const uint64_t key = keys[i];
            if (key_idx(key) != IDX_INVALID && key_dist(key) < bound) c = key_idx(key);
        }
        const uint32_t inv = __ballot_sync(0xFFFFFFFFu, c == IDX_INVALID);
        const uint32_t nvalid = inv ? static_cast<uint32_t>(__ffs(inv) - 1) : 32u;
        unsigned long long sz = 0, lsz = 0;
        if (lane < nvalid) {
            sz = p.walk.offsets[c + 1] - p.walk.offsets[c];
            if (c >= p.walk.list_begin && c < p.walk.list_end) lsz = sz;
        }
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long a = __shfl_up_sync(0xFFFFFFFFu, sz, off), b = __shfl_up_sync(0xFFFFFFFFu, lsz, off);
            if (lane >= static_cast<uint32_t>(off)) { sz += a; lsz += b; }
        }
        const bool cond = lane < nvalid && chosen + lane + 1 >= p.walk.nprobe && reach + sz >= p.walk.k;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, cond);
        const uint32_t take = m ? static_cast<uint32_t>(__ffs(m)) : nvalid;
        if (lane < take) p.walk.probes[q * p.pitch + i] = c;
        if (take > 0) {
            reach += __shfl_sync(0xFFFFFFFFu, sz, take - 1);
            local += __shfl_sync(0xFFFFFFFFu, lsz, take - 1);
        }
        chosen += take;
        if (m) { done = true; break; }
        if (nvalid < 32) break;
    }
    if (lane == 0) {
        if (!done && chosen < p.nlist) atomicExch(p.walk.overflow, 1u);
        p.walk.n_probes[q] = chosen;
        atomicAdd(p.walk.stat_scanned, reach);
        atomicAdd(p.walk.stat_probed, static_cast<unsigned long long>(chosen));
        atomicAdd(p.walk.stat_probed + 1, local);
    }
}

template <int MET>
__global__ void __launch_bounds__(256) coarse_select_kernel(CoarseSelectParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t q = static_cast<uint64_t>(blockIdx.x) * (blockDim.x >> 5) + warp;
    if (q >= p.nq) return;   // warps are independent: no block-wide barrier below
    uint8_t* base = smem + warp * coarse_select_warp_bytes(p.cmax, p.staged_words);
    uint64_t* keys = reinterpret_cast<uint64_t*>(base);                  // [cmax]
    uint32_t* hist = reinterpret_cast<uint32_t*>(keys + p.cmax);         // [256]
    uint32_t* cand = hist + 256;                                          // [cmax]
    uint4* staged = reinterpret_cast<uint4*>(cand + p.cmax);              // [staged_words / 4] ordered images of the row (16-byte aligned: cmax is a multiple of 4)
    // The row of values is either kept in shared memory as ordered integer images (one global pass, coalesced 128-bit loads) or
    // re-read from L2 by every pass (staged_words == 0, the default: occupancy hides the dependent passes better than the
    // staging saves).  dense_ld is a multiple of 128.
    uint32_t thr;
    const uint32_t count = coarse_candidates_full_row(p, q, hist, cand, p.staged_words ? staged : nullptr, lane, thr);
    coarse_emit_ranked<MET>(p, q, thr, count, keys, cand, lane);
}

// The same select from the GROUP MINIMA the dense kernel's epilogue leaves behind (the minimum of every aligned group of 8
// cells, [nq][dense_ld / 8]): the pitch-th smallest group minimum T bounds the pitch-th smallest value from above (pitch
// distinct groups each hold a value <= T), and every cell with a value <= T lies in a group whose minimum is <= T.  So the
// radix select runs over nlist / 8 minima kept in shared memory, and only the selected groups' 8 values each are read from the
// dense matrix -- ~3 KB per query instead of ~5 passes over its 16 KB row at nlist 4096.  The candidate set is again "every
// cell with a value <= thr", so the certificate of coarse_emit_ranked is unchanged.  A query whose selected groups hold more
// than cmax cells at or below thr (cell ids that cluster in space) takes the full-row select instead.
template <int MET>
__global__ void __launch_bounds__(256) coarse_select_gm_kernel(CoarseSelectParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t q = static_cast<uint64_t>(blockIdx.x) * (blockDim.x >> 5) + warp;
    if (q >= p.nq) return;
    uint8_t* base = smem + warp * coarse_gm_warp_bytes(p.cmax, p.dense_ld >> 3, p.c_ld, p.dim);
    uint64_t* keys = reinterpret_cast<uint64_t*>(base);                  // [cmax]; first the selected group ids (uint32 view)
    uint32_t* cand = reinterpret_cast<uint32_t*>(keys + p.cmax);         // [cmax]
    uint32_t* hist = cand + p.cmax;                                       // [256]   (this area is reused for the staged centroid rows)
    uint4* gst4 = reinterpret_cast<uint4*>(hist + 256);                   // [ngrp / 4] ordered images of the group minima
    const uint32_t* gst = reinterpret_cast<const uint32_t*>(gst4);
    uint32_t* glist = reinterpret_cast<uint32_t*>(keys);
    const uint32_t ngrp = p.dense_ld >> 3, ngv = (p.nlist + 7) >> 3, n4 = (ngv + 3) >> 2;   // ngrp is a multiple of 16
    const float4* g4 = reinterpret_cast<const float4*>(p.gmin + q * ngrp);
    uint32_t umin = 0xFFFFFFFFu, umax = 0u;
    for (uint32_t i4 = lane; i4 < n4; i4 += 32) {
        const float4 x = __ldg(g4 + i4);
        uint32_t u[4] = {f32_to_ordered(x.x), f32_to_ordered(x.y), f32_to_ordered(x.z), f32_to_ordered(x.w)};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if (4 * i4 + e < ngv) { umin = min(umin, u[e]); umax = max(umax, u[e]); }
            else u[e] = 0xFFFFFFFFu;
        }
        gst4[i4] = make_uint4(u[0], u[1], u[2], u[3]);
    }
    __syncwarp();
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        umin = min(umin, __shfl_xor_sync(0xFFFFFFFFu, umin, off));
        umax = max(umax, __shfl_xor_sync(0xFFFFFFFFu, umax, off));
    }
    auto fetch = [&](uint32_t i4, uint32_t u[4]) { const uint4 w = gst4[i4]; u[0] = w.x; u[1] = w.y; u[2] = w.z; u[3] = w.w; };
    uint32_t thr = coarse_radix_threshold(fetch, n4, ngv, p.pitch, p.glimit, umin, umax, hist, lane);
    // selected groups, compacted (any order: the candidates are sorted by (distance, cell) later)
    uint32_t n_sel = 0;
    for (uint32_t i0 = 0; i0 < ngv; i0 += 32) {
        const uint32_t g = i0 + lane;
        const bool sel = g < ngv && gst[g] <= thr;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, sel);
        const uint32_t pos = n_sel + __popc(m & ((1u << lane) - 1u));
        if (sel && pos < 2 * p.cmax) glist[pos] = g;     // the keys area holds 2 cmax words
        n_sel += __popc(m);
    }
    __syncwarp();
    uint32_t count = n_sel <= 2 * p.cmax ? 0u : p.cmax + 1u;
    if (count == 0) {
        const float* row = p.dense + q * p.dense_ld;
        for (uint32_t j0 = 0; j0 < n_sel; j0 += 32) {
            const uint32_t j = j0 + lane;
            uint32_t mine = 0, hits = 0, g = 0;
            if (j < n_sel) {
                g = glist[j];
                float4 a, b;
                if (p.blocked) {
                    const float4* pa = reinterpret_cast<const float4*>(p.dense) + dense_piece_index(q, 2 * g, p.dense_ld >> 7);
                    a = __ldg(pa); b = __ldg(pa + 128);      // pieces 2g and 2g + 1 of one column tile sit 128 float4 apart
                } else {
                    a = __ldg(reinterpret_cast<const float4*>(row + 8 * g)); b = __ldg(reinterpret_cast<const float4*>(row + 8 * g) + 1);
                }
                const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int e = 0; e < 8; e++)
                    if (8 * g + e < p.nlist && f32_to_ordered(v[e]) <= thr) { hits |= 1u << e; mine++; }
            }
            uint32_t incl = mine;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
                if (lane >= static_cast<uint32_t>(off)) incl += t;
            }
            uint32_t pos = count + incl - mine;
#pragma unroll
            for (int e = 0; e < 8; e++)
                if (hits & (1u << e)) { if (pos < p.cmax) cand[pos] = 8 * g + e; pos++; }
            count += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        __syncwarp();
    }
    if (count > p.cmax) count = coarse_candidates_full_row(p, q, hist, cand, nullptr, lane, thr);
    coarse_emit_ranked<MET>(p, q, thr, count, keys, cand, lane, reinterpret_cast<uint8_t*>(hist));
}

static __global__ void fill_aux_kernel(float* __restrict__ aux, uint64_t n, uint64_t n_total, float value) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i < n_total) aux[i] = i < n ? value : INFINITY;
}

// ----------------------------------------------------------------------------------------------- build-side assignment
// assign_all_parallel / direct_assign (src/utils/k_means_utils.rs:2119-2241): argmax_c of 2 x.c - |c|^2 (L2) or
// x.c / |c| (cosine), lowest cell on ties.  The flat kernel's selection value IS the negated score (|c|^2 - 2 x.c, or
// -x.c/|c|), so the centroid table is searched like a flat f32 index with k' = 16 and this kernel finishes the job: one
// warp per data row recomputes the score of every retained cell at or below the final threshold in the reference's
// arithmetic, takes the best, and certifies it -- every cell that was not retained has an approximate value >= the
// threshold, so the winner is exact iff its negated score lies below the threshold by more than the error bound.
// Rows that fail are listed for the exact CUDA-core kernel.
struct AssignSelectParams {
    const uint64_t* part_keys;   // [n][parts][kp]
    uint32_t parts, kp;
    const uint32_t* gtau;        // [n] final shared threshold (ordered image of the approximate value)
    uint64_t n;
    const float* rows;           // data rows [n][x_ld]
    uint32_t x_ld;
    const float* centroids;      // [nlist][c_ld]
    uint32_t c_ld;
    const float* aux;            // direct_assign constants: |c|^2 (L2) or 1/|c| (cosine, 0 for a zero centroid)
    uint32_t dim;
    int cosine;
    float eps;
    const uint32_t* cmax_sq_bits; // L2: bit pattern of max |c|^2
    uint32_t* assign_out;        // [n]
    uint32_t* uncert;            // [0] = count, [1..] = row numbers that could not be certified
};

__global__ void __launch_bounds__(256) assign_select_kernel(AssignSelectParams p) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t q = static_cast<uint64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= p.n) return;
    const uint32_t thr = p.gtau[q];
    const uint64_t* src = p.part_keys + q * (static_cast<uint64_t>(p.parts) * p.kp);
    const uint32_t total = p.parts * p.kp;
    const uint8_t* xrow = reinterpret_cast<const uint8_t*>(p.rows + q * p.x_ld);
    float best_s = -INFINITY;
    uint32_t best_c = IDX_INVALID;
    for (uint32_t i = lane; i < total; i += 32) {
        const uint64_t key = src[i];
        if (key == KEY_SENTINEL || static_cast<uint32_t>(key >> 32) > thr) continue;
        const uint32_t c = key_idx(key);
        float dot[1];
        accumulate_fp<4, 4, false, 1>(reinterpret_cast<const uint8_t*>(p.centroids + static_cast<uint64_t>(c) * p.c_ld), xrow, p.x_ld * 4, p.dim, dot);
        const float a = p.aux[c];
        const float score = p.cosine ? __fmul_rn(dot[0], a) : __fsub_rn(__fmul_rn(2.0f, dot[0]), a);
        if (score > best_s || (score == best_s && c < best_c)) { best_s = score; best_c = c; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float os = __shfl_xor_sync(0xFFFFFFFFu, best_s, off);
        const uint32_t oc = __shfl_xor_sync(0xFFFFFFFFu, best_c, off);
        if (os > best_s || (os == best_s && oc < best_c)) { best_s = os; best_c = oc; }
    }
    float qn2 = 0.f;
    for (uint32_t e = lane; e < p.dim; e += 32) {
        const float x = reinterpret_cast<const float*>(xrow)[e];
        qn2 = fmaf(x, x, qn2);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) qn2 += __shfl_xor_sync(0xFFFFFFFFu, qn2, off);
    if (lane == 0) {
        const float tv = ordered_to_f32(thr);   // NaN while no list was ever full: nothing is certified then
        float margin;
        if (p.cosine) {
            margin = p.eps * sqrtf(qn2);                                        // value = -x.c/|c|
        } else {
            const float sn = sqrtf(qn2) + sqrtf(__uint_as_float(*p.cmax_sq_bits));   // value = |c|^2 - 2 x.c
            margin = p.eps * sn * sn;
        }
        const bool ok = best_c != IDX_INVALID && (-best_s) < tv - margin;
        p.assign_out[q] = best_c;
        if (!ok) p.uncert[1 + atomicAdd(p.uncert, 1u)] = static_cast<uint32_t>(q);
    }
}

// tensor-kernel row constants from the direct_assign constants: L2 |c|^2, cosine -1/|c|; +inf on pad rows
static __global__ void assign_tc_aux_kernel(const float* __restrict__ assign_aux, uint32_t nlist, uint64_t total, int cosine, float* __restrict__ aux) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i < total) aux[i] = i < nlist ? (cosine ? -assign_aux[i] : assign_aux[i]) : INFINITY;
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
    float* d_aux2 = nullptr;  // KIND_F16X3: [n_pad + BN] inverse operand scale of every row
    CUtensorMap tm_x;
    DevBuf q_op, part, dbg, gtau, dbgc, dense, dense_gm, q_scale;
    float db_inv_scale = 1.0f;   // KIND_F16X3, L2: 1 / (uniform operand scale)
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
int tc_make_tmap(CUtensorMap* tm, void* base, uint64_t rows, uint32_t kp_elems, int elem_bytes, uint32_t box_rows);
int tc_make_tmap(CUtensorMap* tm, void* base, uint64_t rows, uint32_t kp_elems, int elem_bytes) { return tc_make_tmap(tm, base, rows, kp_elems, elem_bytes, tc::BM); }
int tc_make_tmap(CUtensorMap* tm, void* base, uint64_t rows, uint32_t kp_elems, int elem_bytes, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_last_error("cuTensorMapEncodeTiled is unavailable (driver too old?)"); return ANNB_ERR_CUDA; }
    cuuint64_t dims[2] = {kp_elems, rows};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(kp_elems) * elem_bytes};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(tc::SLAB_BYTES / elem_bytes), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8), 2, base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r))); return ANNB_ERR_CUDA; }
    return ANNB_OK;
}

uint32_t tc_blocks_for(uint64_t work) { return static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>((work + 255) / 256, 148 * 16))); }

// Largest stored-row norm from the per-row constants (L2 only; cosine needs no norm bound).
static int tc_aux_norm_max(annb_index* ix, const float* d_aux, uint64_t n, float* out);
int tc_compute_xnorm_max(annb_index* ix, const float* d_aux, uint64_t n) {
    ix->tc_xnorm_max = 0.f;
    if (ix->metric != ANNB_L2) return ANNB_OK;
    return tc_aux_norm_max(ix, d_aux, n, &ix->tc_xnorm_max);
}
// sqrt of the largest of n squared norms held in d_aux
static int tc_aux_norm_max(annb_index* ix, const float* d_aux, uint64_t n, float* out) {
    uint32_t* d_bits = nullptr;
    ANNB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_bits), 4));
    ANNB_CUDA_CHECK(cudaMemsetAsync(d_bits, 0, 4, ix->stream));
    tc::aux_max_kernel<<<148, 256, 0, ix->stream>>>(d_aux, n, d_bits);
    uint32_t bits = 0;
    cudaError_t e = cudaMemcpyAsync(&bits, d_bits, 4, cudaMemcpyDeviceToHost, ix->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
    cudaFree(d_bits);
    ANNB_CUDA_CHECK(e);
    float sq;
    std::memcpy(&sq, &bits, 4);
    *out = std::sqrt(sq);
    return ANNB_OK;
}

int tc_uniform_f16_scale(annb_index* ix, const float* d_values, uint64_t count, float* out_scale) {
    uint32_t* d_bits = nullptr;
    ANNB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_bits), 4));
    ANNB_CUDA_CHECK(cudaMemsetAsync(d_bits, 0, 4, ix->stream));
    tc::absmax_kernel<<<148 * 8, 256, 0, ix->stream>>>(d_values, count, d_bits);
    uint32_t bits = 0;
    cudaError_t e = cudaMemcpyAsync(&bits, d_bits, 4, cudaMemcpyDeviceToHost, ix->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
    cudaFree(d_bits);
    ANNB_CUDA_CHECK(e);
    float m;
    std::memcpy(&m, &bits, 4);
    int ex = 0, sh = 0;
    if (m > 0.f) { (void)std::frexp(m, &ex); sh = std::min(std::max(14 - ex, -100), 100); }   // m = f 2^ex, f in [0.5, 1): m 2^sh in [2^13, 2^14)
    *out_scale = std::ldexp(1.0f, sh);
    return ANNB_OK;
}

int tc_flat_prepare(annb_index* ix) {
    if (ix->is_ivf) return ANNB_OK;
    int kind = ix->dtype == ANNB_F32 ? tc::KIND_TF32X3 : (ix->dtype == ANNB_BF16 ? tc::KIND_BF16 : tc::KIND_I8);
    // f32 rows of up to 256 elements: 3xFP16 (rows scaled by powers of two, split_f16_kernel) -- the three-term product of 3xTF32
    // at 16 instead of 8 elements per MMA K step.  Option tc_f32_fp16 = 0 keeps 3xTF32; wider rows keep it as well.
    if (kind == tc::KIND_TF32X3 && ix->opt_tc_f32_fp16 != 0 && round_up(ix->dim, 64u) <= 256u) kind = tc::KIND_F16X3;
    const uint32_t elem = kind == tc::KIND_TF32X3 ? 4 : (kind == tc::KIND_I8 ? 1 : 2);
    const uint32_t slab_elems = tc::SLAB_BYTES / elem;
    const uint32_t kp = round_up(ix->dim, slab_elems);
    // the query tile lives in TMEM: 512 B of a row per piece (f32 dim <= 128, bf16 <= 256, int8 <= 512); f32 rows of up to 1024 B
    // keep only their hi piece there and the lo piece in shared memory (Params::lo_smem).  Wider rows: the exact CUDA-core path.
    // Rows of up to 2048 B (f32 dim <= 512) stream their query slabs through the ring with the database (Params::stream_q).
    if (kp * elem > 2048u) return ANNB_OK;
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
    tc::aux_kernel<<<static_cast<uint32_t>((aux_rows + 127) / 128), 128, 0, s>>>(ix->d_rows, ix->row_bytes, kind, ix->dim, ix->d_norms, ix->d_norms_i,
                                                                                ix->metric == ANNB_COSINE, ix->n, aux_rows, st->d_aux);
    ANNB_CUDA_CHECK(cudaGetLastError());
    void* xbase = nullptr;
    uint64_t xrows = 0;
    if (kind == tc::KIND_F16X3) {
        const uint64_t bytes = 2ull * st->n_pad * kp * 2;
        cudaError_t e = cudaMalloc(&st->d_x, bytes);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&st->d_aux2), aux_rows * sizeof(float));
        if (e != cudaSuccess) { (void)cudaGetLastError(); set_last_error(std::string("cudaMalloc tc operand: ") + cudaGetErrorString(e)); return ANNB_ERR_OUT_OF_MEMORY; }
        st->bytes += bytes + aux_rows * sizeof(float);
        tc::fill_aux_kernel<<<static_cast<uint32_t>((aux_rows + 127) / 128), 128, 0, s>>>(st->d_aux2, aux_rows, aux_rows, 1.0f);
        if (ix->metric == ANNB_COSINE)     // cosine: unit rows at a uniform scale, no per-row constant at all (flat_tc_kernel, UNIT; the aux arrays stay unused)
            tc::split_f16_unit_kernel<<<tc_blocks_for(static_cast<uint64_t>(st->n_pad) * 32), 256, 0, s>>>(reinterpret_cast<const float*>(ix->d_rows), ix->row_bytes / 4, ix->dim,
                                                                                                     ix->n, st->n_pad, kp, ix->d_norms, static_cast<__half*>(st->d_x));
        else {                             // L2 keeps |x|^2 as its per-row constant; one uniform scale for the whole operand
            float sc = 1.0f;
            ANNB_TRY(tc_uniform_f16_scale(ix, reinterpret_cast<const float*>(ix->d_rows), ix->n * static_cast<uint64_t>(ix->row_bytes / 4), &sc));
            st->db_inv_scale = 1.0f / sc;
            tc::split_f16_kernel<<<tc_blocks_for(static_cast<uint64_t>(st->n_pad) * 32), 256, 0, s>>>(reinterpret_cast<const float*>(ix->d_rows), ix->row_bytes / 4, ix->dim, ix->n,
                                                                                                st->n_pad, kp, static_cast<__half*>(st->d_x), nullptr, 1.0f, sc);
        }
        ANNB_CUDA_CHECK(cudaGetLastError());
        xbase = st->d_x;
        xrows = 2ull * st->n_pad;
    } else if (kind == tc::KIND_TF32X3) {
        const uint64_t bytes = 2ull * st->n_pad * kp * 4;
        cudaError_t e = cudaMalloc(&st->d_x, bytes);
        if (e != cudaSuccess) { (void)cudaGetLastError(); set_last_error(std::string("cudaMalloc tc operand: ") + cudaGetErrorString(e)); return ANNB_ERR_OUT_OF_MEMORY; }
        st->bytes += bytes;
        tc::split_tf32_kernel<<<tc_blocks_for(static_cast<uint64_t>(st->n_pad) * kp), 256, 0, s>>>(reinterpret_cast<const float*>(ix->d_rows), ix->row_bytes / 4,
                                                                                                ix->dim, ix->n, st->n_pad, kp, static_cast<float*>(st->d_x));
        ANNB_CUDA_CHECK(cudaGetLastError());
        xbase = st->d_x;
        xrows = 2ull * st->n_pad;
    } else if (kind == tc::KIND_I8) {
        const uint64_t bytes = static_cast<uint64_t>(st->n_pad) * kp;
        cudaError_t e = cudaMalloc(&st->d_x, bytes);
        if (e != cudaSuccess) { (void)cudaGetLastError(); set_last_error(std::string("cudaMalloc tc operand: ") + cudaGetErrorString(e)); return ANNB_ERR_OUT_OF_MEMORY; }
        st->bytes += bytes;
        tc::pad_i8_kernel<<<tc_blocks_for(static_cast<uint64_t>(st->n_pad) * kp), 256, 0, s>>>(reinterpret_cast<const int8_t*>(ix->d_rows), ix->row_bytes, ix->dim, ix->n,
                                                                                            st->n_pad, kp, static_cast<int8_t*>(st->d_x));
        ANNB_CUDA_CHECK(cudaGetLastError());
        xbase = st->d_x;
        xrows = st->n_pad;
    } else {
        const uint64_t bytes = static_cast<uint64_t>(st->n_pad) * kp * 2;
        cudaError_t e = cudaMalloc(&st->d_x, bytes);
        if (e != cudaSuccess) { (void)cudaGetLastError(); set_last_error(std::string("cudaMalloc tc operand: ") + cudaGetErrorString(e)); return ANNB_ERR_OUT_OF_MEMORY; }
        st->bytes += bytes;
        tc::pad_bf16_kernel<<<tc_blocks_for(static_cast<uint64_t>(st->n_pad) * kp), 256, 0, s>>>(reinterpret_cast<const uint16_t*>(ix->d_rows), ix->row_bytes / 2,
                                                                                              ix->dim, ix->n, st->n_pad, kp, static_cast<uint16_t*>(st->d_x));
        ANNB_CUDA_CHECK(cudaGetLastError());
        xbase = st->d_x;
        xrows = st->n_pad;
    }
    ANNB_TRY(tc_make_tmap(&st->tm_x, xbase, xrows, kp, elem));
    ANNB_TRY(tc_compute_xnorm_max(ix, st->d_aux, ix->n));
    ix->device_bytes += st->bytes;
    return ANNB_OK;
}

void tc_destroy(annb_index* ix) {
    if (!ix->tc) return;
    ix->device_bytes -= std::min<uint64_t>(ix->device_bytes, ix->tc->bytes);
    cudaFree(ix->tc->d_x);
    cudaFree(ix->tc->d_aux);
    cudaFree(ix->tc->d_aux2);
    ix->tc->q_scale.release();
    ix->tc->q_op.release();
    ix->tc->part.release();
    ix->tc->dbg.release();
    ix->tc->gtau.release();
    ix->tc->dbgc.release();
    delete ix->tc;
    ix->tc = nullptr;
}

// Error bound of the tensor-core pre-selection values, in the units the coverage certificates test:
// |reference-order distance - distance implied by the selection value| <= eps * (|q| + |x|max)^2 (L2) or eps (cosine).
//   E  = bound of |s - q.x| / (|q||x|) for the accumulated tensor-core dot s:
//        operand representation (f32: dropped lo.lo term + residuals of the hi/lo splits; bf16 index: residual of the
//        query's bf16 terms, the stored rows are exact) + TC_MMA_ULPS * 2^-23 per accumulating MMA instruction of a tile
//        row (each tcgen05.mma K step adds its products to the f32 accumulator; modelled as at most TC_MMA_ULPS ulp of
//        a partial sum bounded by |q||x| -- truncation or better; tests/test_gpu_tensor.py::test_adversarial_* measures it)
//   +  the reference's own rounding: 8 lanes x dim/8 sequential adds + the reduce tree + finish (dist.rs:306-330)
//   +  the epilogue's fma / multiply, the row constant (computed in f64, rounded once) and the value -> distance map.
constexpr double TC_MMA_ULPS = 1.0;
float tc_cert_eps(const annb_index* ix, int kind, uint32_t kp_elems, uint32_t terms, bool inkernel_split) {
    if (ix->opt_cert_eps >= 0.f) return ix->opt_cert_eps;     // caller override ("cert_eps_log2"), 0 = certificate off
    const bool l2 = ix->metric == ANNB_L2;
    const double u23 = std::ldexp(1.0, -23), u24 = std::ldexp(1.0, -24);
    double E;
    if (kind == tc::KIND_I8) {
        // exact integer dots.  L2 (dim <= 256: sums < 2^24 are exact in f32): only a tie between the k-th distance and the
        // k'-th selection value has to be excluded (pruning is strict); cosine: a few roundings on an exact dot.
        if (l2 && ix->dim <= 256) return 1e-30f;
        if (!l2) return 4.7683716e-07f;   // 2^-21: the reference divides the exact integer dot by two square roots -- a handful of roundings on each side
        E = 0.0;
    } else if (kind == tc::KIND_F16X3) {
        // same split as 3xTF32 (11 significant bits per piece, round to nearest both times; the power-of-two row scales are exact),
        // half the accumulating MMAs per row (16 elements per K step)
        const double n_mma = 3.0 * (kp_elems / 16);
        E = 3.0 * std::ldexp(1.0, -22) + TC_MMA_ULPS * n_mma * u23;
        if (!l2 && !ix->is_ivf) E += 2.0 * u24;     // flat cosine operand: rows divided by their norm in f32 (split_f16_unit_kernel)
    } else if (kind == tc::KIND_TF32X3) {
        const double n_mma = 3.0 * (kp_elems / 8);
        // flat: q and x both split with cvt.rna (2^-22 residual each) + dropped lo.lo (2^-22); IVF in-kernel split: x hi
        // truncated, lo rounded (2^-21 residual), dropped lo.lo <= 2^-21, query residual 2^-22
        E = (inkernel_split ? 5.0 : 3.0) * std::ldexp(1.0, -22) + TC_MMA_ULPS * n_mma * u23;
    } else {
        const double n_mma = static_cast<double>(terms) * (kp_elems / 16);
        // residual of the f32 query after its bf16 terms (each term keeps 8 significant bits, round to nearest: 2^-8 of what
        // it rounds): 2^-16 |q| after two terms, 2^-24 after three; bf16 self queries are exact
        const double repr = terms >= 3 ? std::ldexp(1.0, -24) : (terms == 2 ? std::ldexp(1.0, -16) : 0.0);
        E = repr + TC_MMA_ULPS * n_mma * u23;
    }
    const double ref = (ix->dim / 8 + 8) * u24;
    const double eps = l2 ? (E / 2 + ref) : (E * (1.0 + 1.0 / 128) + ref + 6 * u24);
    return static_cast<float>(eps);
}

// bf16 terms of an f32 query against a BF16 index.  Two terms leave a residual of 2^-16 |q|: the certificate then needs the
// k-th neighbour to lie ~1e-5 (|q| + |x|max)^2 below the pruning threshold.  Cosine distances of a BF16 index are spread by
// the index's own rounding (the reference divides by the norm of the UN-rounded row), two terms certify there; squared
// distances of close neighbours are not (measured on the bench data, tools/gap_stats.py: a quarter of the queries would
// go to the exact fallback), so L2 keeps three terms unless the caller insists.
uint32_t tc_bf16_terms(const annb_index* ix) {
    if (ix->opt_tc_bf16_terms == 2 || ix->opt_tc_bf16_terms == 3) return static_cast<uint32_t>(ix->opt_tc_bf16_terms);
    return ix->metric == ANNB_COSINE ? 2u : 3u;
}

// k' = 16 serves k <= 16: a query whose 16th merged value is too close to its k-th distance takes the re-rank's second chance
// (up to 64 candidates below the scan's final threshold), and a handle whose batches still fail (tc_escalate, see run_batch)
// moves to k' = 32 and then to wide-k mode.  Measured, self-kNN 2M x 50 k = 15: 12.1 ms (k' = 16) against 12.7 ms (k' = 32) per
// 10 000 rows on one GPU, 1.94 against 2.62 ms on a 250k-row shard; no uncertified query either way.
static uint32_t pick_kprime(const annb_index* ix, uint32_t k_eff) {
    if (ix->opt_tc_candidates == 16 || ix->opt_tc_candidates == 32) return std::max<uint32_t>(ix->opt_tc_candidates, k_eff <= 16 ? 16 : 32);
    return (k_eff <= 16 && ix->tc_escalate == 0) ? 16 : 32;
}

// k above TC_K_LIST (one k' = 32 list per thread covers it) and up to TC_K_WIDE: "wide-k" mode.  Every (split, half tile) list
// keeps its own 32 best and prunes with its own threshold only; the lists interleave over the database (strided tiles), so the
// true top-k of a query spread over all of them and the union of >= max(8, k / 8) lists holds it with room to spare.  Whether
// it did is not assumed but checked: the re-rank recomputes every candidate at or below G = the smallest final list threshold
// and certifies the query iff its k-th exact distance lies below the bound of G (rerank_kernel, wide branch); a database whose
// order defeats the interleaving (a query's neighbours all in one 64-row half tile stride) sends those queries to the exact path.
constexpr uint32_t TC_K_LIST = 24, TC_K_WIDE = 256;
static uint32_t wide_k_lists(uint32_t k_eff) { return std::max<uint32_t>(8u, (k_eff + 7u) / 8u); }

// operand form of the flat tensor path: -1 none, 0 3xTF32, 1 bf16 terms, 2 int8, 3 3xFP16
int tc_flat_kind(const annb_index* ix) { return ix->tc ? ix->tc->kind : -1; }

bool tc_flat_supported(const annb_index* ix, int qt, uint32_t k_eff) {
    if (!ix->tc) return false;
    if (k_eff > TC_K_WIDE) return false;
    if (k_eff > TC_K_LIST && (ix->opt_tc_wide_k == 0 || static_cast<uint64_t>(ix->tc->n_pad / tc::BN) * 2 < wide_k_lists(k_eff))) return false;
    if (ix->dtype == ANNB_F32) return qt == QT_F32;
    if (ix->dtype == ANNB_BF16) return qt == QT_F32 || qt == QT_BF16;
    if (ix->dtype == ANNB_SQ8) return qt == QT_I8;
    return false;
}

// Database splits per query tile: fill whole waves of 148 CTAs.
static uint32_t pick_splits(uint64_t q_tiles, uint64_t n_tiles_db, int requested, uint32_t overhead_tiles = 8) {
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
        // work per CTA shrinks with more splits; fixed per-CTA cost ~ 8 tiles' worth (query load, pipeline fill, warm-up of the
        // k' list) for k' = 16.  k' = 32 lists warm up much more slowly and every extra split weakens the shared threshold:
        // measured on a 250k x 50, k = 15 shard, 5 splits 2.74 ms, 9 splits 2.88 ms, 15 splits 3.50 ms -> ~40 tiles' worth.
        const double eff = (static_cast<double>(ctas) / static_cast<double>(waves * sms)) *
                           (static_cast<double>(tiles_per) / static_cast<double>(tiles_per + overhead_tiles));
        if (eff > best_eff + 1e-9) { best_eff = eff; best = static_cast<uint32_t>(splits); }
    }
    return best;
}

template <int KIND, int KP, int MET, bool TS, int EW = 2>
static int launch_tc(const CUtensorMap& tmq, const CUtensorMap& tmx, const tc::Params& p, dim3 grid, size_t smem, cudaStream_t s) {
    auto kern = tc::flat_tc_kernel<KIND, KP, MET, TS, false, EW>;
    ANNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, 64 + EW * 128, smem, s>>>(tmq, tmx, p);
    ANNB_CUDA_CHECK(cudaGetLastError());
    return ANNB_OK;
}

template <int RT, int QT, int MET>
static int launch_rerank(const tc::RerankParams& r, cudaStream_t s) {
    auto kern = tc::rerank_kernel<RT, QT, MET>;
    const size_t smem = (static_cast<size_t>(r.nsort) + (r.n_exact > 64u ? r.n_exact : 0u)) * 8;
    ANNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<static_cast<uint32_t>(r.nq), 128, smem, s>>>(r);
    ANNB_CUDA_CHECK(cudaGetLastError());
    return ANNB_OK;
}

int tc_flat_search(annb_index* ix, const uint8_t* d_q, uint32_t q_bytes, int qt, int bf16_self, uint64_t nq, uint32_t k_eff, uint32_t k_out,
                   uint64_t* d_ids, float* d_dist, uint32_t* d_cnt, cudaStream_t s) {
    TcState* st = ix->tc;
    const int kind = st->kind;
    const uint32_t elem = kind == tc::KIND_TF32X3 ? 4 : (kind == tc::KIND_I8 ? 1 : 2);
    const uint32_t kp = st->kp_elems;
    const bool split3 = kind == tc::KIND_TF32X3 || kind == tc::KIND_F16X3;
    uint32_t kprime = pick_kprime(ix, k_eff);
    // f32 rows of more than 128 elements: the accumulation error of the tensor core grows with the MMAs per tile row (tc_cert_eps),
    // and a k' = 16 list's threshold then sits inside the certificate's margin of the k-th distance on data with ~1e-6 neighbour
    // gaps (measured, 1M x 256 Correlated cosine k = 10: 7.5 % of the queries uncertified with k' = 16)
    if (kind == tc::KIND_TF32X3 && kp > 128 && ix->opt_tc_candidates == 0) kprime = 32;
    if (k_eff > TC_K_LIST) kprime = 32;
    const uint32_t nq_pad = static_cast<uint32_t>(round_up<uint64_t>(nq, tc::BM));
    // f32 queries against a BF16 index go in as two or three bf16 terms q0 + q1 [+ q2]: 16 or 24 MMAs per tile
    const uint32_t bf16_terms = tc_bf16_terms(ix);
    const uint32_t na = split3 ? 2 : ((qt == QT_BF16 || kind == tc::KIND_I8) ? 1 : bf16_terms);

    // ---- query operand: stacked pieces, zero padded ----
    ANNB_TRY(st->q_op.ensure(static_cast<uint64_t>(split3 ? 2 : 3) * nq_pad * kp * elem));
    if (kind == tc::KIND_F16X3) {
        ANNB_TRY(st->q_scale.ensure(static_cast<uint64_t>(nq_pad) * 4));
        tc::split_f16_kernel<<<tc_blocks_for(static_cast<uint64_t>(nq_pad) * 32), 256, 0, s>>>(reinterpret_cast<const float*>(d_q), q_bytes / 4, ix->dim, nq, nq_pad, kp,
                                                                                            st->q_op.as<__half>(), st->q_scale.as<float>(),
                                                                                            ix->metric == ANNB_COSINE ? -1.0f : 1.0f);   // cosine (UNIT): the accumulator is -q.x/|x| itself
    } else if (kind == tc::KIND_I8) {
        tc::pad_i8_kernel<<<tc_blocks_for(static_cast<uint64_t>(nq_pad) * kp), 256, 0, s>>>(reinterpret_cast<const int8_t*>(d_q), q_bytes, ix->dim, nq, nq_pad, kp,
                                                                                         st->q_op.as<int8_t>());
    } else if (kind == tc::KIND_TF32X3) {
        tc::split_tf32_kernel<<<tc_blocks_for(static_cast<uint64_t>(nq_pad) * kp), 256, 0, s>>>(reinterpret_cast<const float*>(d_q), q_bytes / 4, ix->dim, nq, nq_pad, kp,
                                                                                             st->q_op.as<float>());
    } else if (qt == QT_F32) {
        tc::split_bf16x3_kernel<<<tc_blocks_for(static_cast<uint64_t>(nq_pad) * kp), 256, 0, s>>>(reinterpret_cast<const float*>(d_q), q_bytes / 4, ix->dim, nq, nq_pad, kp,
                                                                                               st->q_op.as<__nv_bfloat16>());
    } else {
        tc::pad_bf16_kernel<<<tc_blocks_for(static_cast<uint64_t>(nq_pad) * kp), 256, 0, s>>>(reinterpret_cast<const uint16_t*>(d_q), q_bytes / 2, ix->dim, nq, nq_pad, kp,
                                                                                           st->q_op.as<uint16_t>());
    }
    ANNB_CUDA_CHECK(cudaGetLastError());
    ix->stat_launches++;
    CUtensorMap tmq;
    ANNB_TRY(tc_make_tmap(&tmq, st->q_op.p, static_cast<uint64_t>(na) * nq_pad, kp, elem));

    // ---- geometry ----
    const uint64_t q_tiles = nq_pad / tc::BM;
    const uint64_t db_tiles = st->n_pad / tc::BN;
    // A handle whose batches keep failing the certificate (more than 2 % of a batch went to the exact fallback: run_batch raises
    // tc_escalate -- first to k' = 32, then to this) switches to the wide-k machinery whatever k is: the union of >= 8 unshared k' = 32 lists reaches far beyond
    // the k-th neighbour, which a wide certificate margin needs (1M x 256 Correlated cosine, k = 10: 748 uncertified queries per
    // 10 000 with a shared threshold and the 64-candidate second chance -- 65 ms per batch; 0 in wide mode -- 27 ms).
    const bool wide_k = k_eff > TC_K_LIST || (ix->tc_escalate >= 2 && ix->opt_tc_wide_k != 0);
    uint32_t splits_req = pick_splits(q_tiles, db_tiles, ix->opt_db_splits, kprime == 32 ? 40u : 8u);
    if (wide_k) splits_req = static_cast<uint32_t>(std::min<uint64_t>(db_tiles, std::max<uint32_t>(splits_req, (wide_k_lists(k_eff) + 1) / 2)));
    const uint64_t tiles_per = (db_tiles + splits_req - 1) / splits_req;
    const uint32_t splits = wide_k ? splits_req : static_cast<uint32_t>((db_tiles + tiles_per - 1) / tiles_per);
    const uint32_t nb = split3 ? 2 : 1;
    // query operand resident in TMEM (TS-mode MMA) when its pieces fit their column budget (bf16 terms: 64 columns each)
    const uint32_t bf16_piece_cols = kp > 128 ? 128u : 64u;
    const bool ts = ix->opt_tc_ts != 0 && (kind != tc::KIND_BF16 || na * bf16_piece_cols <= 256) && kp * elem <= (kind == tc::KIND_TF32X3 ? 1024u : 512u);
    const bool hyb = ts && kind == tc::KIND_BF16 && na == 3 && ix->opt_tc_bf16_hybrid != 0;
    // f32 rows of 129 .. 256 elements: hi piece in TMEM, lo piece in shared memory (option tc_f32_lo_smem = 1 forces it for narrow rows too)
    const bool lo_s = ts && split3 && ((kind == tc::KIND_TF32X3 && kp > 128) || ix->opt_tc_f32_lo_smem != 0);
    const size_t fixed = 256 /*barriers*/ + 2 * 8 * 64 * 4 /*per-warp row constants, two banks*/;
    const size_t budget = 227 * 1024;
    size_t q_smem = ts ? ((hyb || lo_s) ? static_cast<size_t>(st->nslab) * tc::SLAB_TILE : 0) : static_cast<size_t>(split3 ? 2 : (kind == tc::KIND_I8 ? 1 : 3)) * st->nslab * tc::SLAB_TILE;
    // SS mode with a query tile that leaves fewer than two ring stages: stream the query slabs with the database slabs instead
    const bool stream_q = !ts && q_smem + fixed + 2 * nb * tc::SLAB_TILE > budget;
    if (stream_q) q_smem = 0;
    const uint32_t stage_slabs = nb + (stream_q ? na : 0u);
    uint32_t stages = static_cast<uint32_t>((budget - q_smem - fixed) / (stage_slabs * tc::SLAB_TILE));
    stages = std::min<uint32_t>(stages, 8);
    const size_t smem = q_smem + static_cast<size_t>(stages) * stage_slabs * tc::SLAB_TILE + fixed;

    // epilogue warps per TMEM lane quarter: the int8 and bf16 kernels (few MMAs per tile, epilogue-paced) run four, k' = 16 only
    // (a k' = 32 list does not fit the 112-register budget of 18 warps); option tc_epi_warps = 2 forces the narrow layout
    // (measured: four epilogue warps per quarter are SLOWER -- bf16 4.84 -> 5.29 ms, int8 3.15 -> 3.80 ms on 1M x 128: the epilogue
    // is bound by the 64 B / cycle TMEM read path, not by latency; kept behind option tc_epi_warps = 4)
    const uint32_t ew = (!split3 && ts && kprime == 16 && ix->opt_tc_epi_warps == 4) ? 4u : 2u;
    ANNB_TRY(st->part.ensure(nq * ew * splits * static_cast<uint64_t>(kprime) * 8));
    ANNB_TRY(st->gtau.ensure(static_cast<uint64_t>(nq_pad) * 4));
    ANNB_CUDA_CHECK(cudaMemsetAsync(st->gtau.p, 0xFF, static_cast<uint64_t>(nq_pad) * 4, s));
    tc::Params p{};
    p.nq = nq; p.n_rows = ix->n; p.nq_pad = nq_pad; p.n_pad = st->n_pad; p.nslab = st->nslab; p.n_stages = stages;
    p.n_splits = splits; p.rows_per_split = tiles_per * tc::BN; p.a_pieces = na; p.aux = st->d_aux; p.hybrid = hyb ? 1u : 0u;
    p.aux2 = st->d_aux2; p.q_inv_scale = st->q_scale.as<float>(); p.db_inv_scale = st->db_inv_scale;
    p.lo_smem = lo_s ? 1u : 0u; p.stream_q = stream_q ? 1u : 0u; p.wide_k = wide_k ? 1u : 0u; p.strided = (wide_k || ix->opt_tc_strided != 0) ? 1u : 0u;
    p.part_keys = st->part.as<uint64_t>(); p.dbg = st->dbg.as<float>(); p.gtau = st->gtau.as<uint32_t>(); p.q_op = st->q_op.as<void>(); p.kp = kp; p.dbg_cycles = st->dbgc.as<unsigned long long>();
    {
        dim3 grid(static_cast<uint32_t>(q_tiles), splits);
        // timed as the dominant kernel of the flat path
        cudaEvent_t ea = nullptr, eb = nullptr;
        if (ix->opt_time_kernels && cudaEventCreate(&ea) == cudaSuccess && cudaEventCreate(&eb) == cudaSuccess) cudaEventRecord(ea, s);
        int rc;
        const bool l2 = ix->metric == ANNB_L2;
#define ANNB_TC_LAUNCH(KIND_, KP_, TS_) (l2 ? launch_tc<KIND_, KP_, MET_L2, TS_>(tmq, st->tm_x, p, grid, smem, s) : launch_tc<KIND_, KP_, MET_COS, TS_>(tmq, st->tm_x, p, grid, smem, s))
#define ANNB_TC_LAUNCH4(KIND_) (l2 ? launch_tc<KIND_, 16, MET_L2, true, 4>(tmq, st->tm_x, p, grid, smem, s) : launch_tc<KIND_, 16, MET_COS, true, 4>(tmq, st->tm_x, p, grid, smem, s))
        if (ew == 4) rc = kind == tc::KIND_I8 ? ANNB_TC_LAUNCH4(tc::KIND_I8) : ANNB_TC_LAUNCH4(tc::KIND_BF16);
        else if (kind == tc::KIND_I8 && ts) rc = kprime == 16 ? ANNB_TC_LAUNCH(tc::KIND_I8, 16, true) : ANNB_TC_LAUNCH(tc::KIND_I8, 32, true);
        else if (kind == tc::KIND_I8) rc = kprime == 16 ? ANNB_TC_LAUNCH(tc::KIND_I8, 16, false) : ANNB_TC_LAUNCH(tc::KIND_I8, 32, false);
        else if (kind == tc::KIND_F16X3 && ts) rc = kprime == 16 ? ANNB_TC_LAUNCH(tc::KIND_F16X3, 16, true) : ANNB_TC_LAUNCH(tc::KIND_F16X3, 32, true);
        else if (kind == tc::KIND_F16X3) rc = kprime == 16 ? ANNB_TC_LAUNCH(tc::KIND_F16X3, 16, false) : ANNB_TC_LAUNCH(tc::KIND_F16X3, 32, false);
        else if (kind == tc::KIND_TF32X3 && ts) rc = kprime == 16 ? ANNB_TC_LAUNCH(tc::KIND_TF32X3, 16, true) : ANNB_TC_LAUNCH(tc::KIND_TF32X3, 32, true);
        else if (kind == tc::KIND_TF32X3) rc = kprime == 16 ? ANNB_TC_LAUNCH(tc::KIND_TF32X3, 16, false) : ANNB_TC_LAUNCH(tc::KIND_TF32X3, 32, false);
        else if (ts) rc = kprime == 16 ? ANNB_TC_LAUNCH(tc::KIND_BF16, 16, true) : ANNB_TC_LAUNCH(tc::KIND_BF16, 32, true);
        else rc = kprime == 16 ? ANNB_TC_LAUNCH(tc::KIND_BF16, 16, false) : ANNB_TC_LAUNCH(tc::KIND_BF16, 32, false);
#undef ANNB_TC_LAUNCH
#undef ANNB_TC_LAUNCH4
        if (ea && eb) { cudaEventRecord(eb, s); ix->timed.emplace_back(ea, eb); }
        ANNB_TRY(rc);
        ix->stat_launches++;
    }
    // ---- exact re-rank + merge ----
    tc::RerankParams r{};
    r.part_keys = st->part.as<uint64_t>(); r.parts = ew * splits; r.kp = kprime; r.k_eff = k_eff; r.k_out = k_out; r.gtau = st->gtau.as<uint32_t>();
    r.nsort = next_pow2(std::max(ew * splits * kprime, 64u));
    r.n_exact = wide_k ? std::min<uint32_t>(r.nsort, 2048u) : 0u;     // every candidate below the certificate's threshold, up to 2048 per query
    if (wide_k) r.n_exact = std::max<uint32_t>(r.n_exact, 128u);
    r.nq = nq; r.rows = ix->d_rows; r.row_bytes = ix->row_bytes; r.row_norms = ix->d_norms; r.row_norms_i = ix->d_norms_i; r.queries = d_q; r.q_bytes = q_bytes; r.dim = ix->dim;
    r.bf16_self = bf16_self; r.id_base = ix->id_base; r.out_ids = d_ids; r.out_dist = d_dist; r.out_counts = d_cnt;
    ANNB_TRY(ix->s_uncert.ensure((nq + 1) * 4));
    ANNB_CUDA_CHECK(cudaMemsetAsync(ix->s_uncert.p, 0, 4, s));
    r.cert_eps = tc_cert_eps(ix, kind, kp, na, false);
    { uint32_t b; std::memcpy(&b, &r.cert_eps, 4); ix->stat_cert_eps_bits = b; }
    r.xnorm_max = ix->tc_xnorm_max; r.uncert_count = ix->s_uncert.as<uint32_t>(); r.uncert_list = ix->s_uncert.as<uint32_t>() + 1; r.out_bound = ix->shard_bound;
    const bool cos = ix->metric == ANNB_COSINE;
    int rc;
    if (ix->dtype == ANNB_SQ8) rc = cos ? launch_rerank<2, QT_I8, MET_COS>(r, s) : launch_rerank<2, QT_I8, MET_L2>(r, s);
    else if (ix->dtype == ANNB_F32) rc = cos ? launch_rerank<0, QT_F32, MET_COS>(r, s) : launch_rerank<0, QT_F32, MET_L2>(r, s);
    else if (qt == QT_F32) rc = cos ? launch_rerank<1, QT_F32, MET_COS>(r, s) : launch_rerank<1, QT_F32, MET_L2>(r, s);
    else rc = cos ? launch_rerank<1, QT_BF16, MET_COS>(r, s) : launch_rerank<1, QT_BF16, MET_L2>(r, s);
    ANNB_TRY(rc);
    ix->stat_launches++;
    return ANNB_OK;
}

// ----------------------------------------------------------------------------------------------- IVF centroid ranking
// The centroid table as a tensor-core operand (3xTF32, like a flat f32 index); see coarse_select_kernel.
int tc_coarse_prepare(annb_index* ix) {
    if (!ix->is_ivf || ix->nlist < 512) return ANNB_OK;      // small tables: the exact CUDA-core ranking is already cheap
    // 3xFP16 (rows scaled by powers of two, 16 elements per MMA K step) like the flat f32 index; option ivf_coarse_fp16 = 0
    // rebuilds this state as 3xTF32
    const bool f16 = ix->opt_ivf_coarse_fp16 != 0;
    const uint32_t elem = f16 ? 2u : 4u;
    const uint32_t slab_elems = tc::SLAB_BYTES / elem;
    const uint32_t kp = round_up(ix->dim, slab_elems);
    if (round_up(ix->dim, 32u) * 4 > 512) return ANNB_OK;
    TcState* st = new TcState();
    ix->tc_coarse = st;
    st->kind = f16 ? tc::KIND_F16X3 : tc::KIND_TF32X3;
    st->kp_elems = kp;
    st->nslab = kp / slab_elems;
    st->n_pad = round_up<uint32_t>(ix->nlist, tc::BN);
    cudaStream_t s = ix->stream;
    const uint64_t aux_rows = static_cast<uint64_t>(st->n_pad) + tc::BN;
    ANNB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&st->d_aux), aux_rows * sizeof(float)));
    const uint32_t ag = static_cast<uint32_t>((aux_rows + 127) / 128);
    const uint8_t* crow = reinterpret_cast<const uint8_t*>(ix->d_centroids);
    tc::aux_kernel<<<ag, 128, 0, s>>>(crow, ix->cent_ld * 4, 0, ix->dim, nullptr, nullptr, 0, ix->nlist, aux_rows, st->d_aux);
    ANNB_CUDA_CHECK(cudaGetLastError());
    ANNB_TRY(tc_aux_norm_max(ix, st->d_aux, ix->nlist, &ix->tc_cnorm_max));
    if (ix->metric == ANNB_COSINE) {
        if (ix->dtype == ANNB_SQ8) tc::fill_aux_kernel<<<ag, 128, 0, s>>>(st->d_aux, ix->nlist, aux_rows, -1.0f);   // pre-normalised: value = -q.c
        else tc::aux_kernel<<<ag, 128, 0, s>>>(crow, ix->cent_ld * 4, 0, ix->dim, ix->d_centroid_norms, nullptr, 1, ix->nlist, aux_rows, st->d_aux);
        ANNB_CUDA_CHECK(cudaGetLastError());
    }
    const uint64_t bytes = 2ull * st->n_pad * kp * elem;
    ANNB_CUDA_CHECK(cudaMalloc(&st->d_x, bytes));
    st->bytes = bytes + aux_rows * sizeof(float);
    if (f16) {
        ANNB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&st->d_aux2), aux_rows * sizeof(float)));
        st->bytes += aux_rows * sizeof(float);
        tc::fill_aux_kernel<<<ag, 128, 0, s>>>(st->d_aux2, aux_rows, aux_rows, 1.0f);
        float sc = 0.0f;                   // L2: one uniform scale for the table (0 = per-row scales, cosine)
        if (ix->metric != ANNB_COSINE) {
            ANNB_TRY(tc_uniform_f16_scale(ix, ix->d_centroids, static_cast<uint64_t>(ix->nlist) * ix->cent_ld, &sc));
            st->db_inv_scale = 1.0f / sc;
        }
        tc::split_f16_kernel<<<tc_blocks_for(static_cast<uint64_t>(st->n_pad) * 32), 256, 0, s>>>(ix->d_centroids, ix->cent_ld, ix->dim, ix->nlist, st->n_pad, kp,
                                                                                            static_cast<__half*>(st->d_x), st->d_aux2, 1.0f, sc);
        if (ix->metric == ANNB_COSINE)     // cosine: one constant per cell, -1 / (|c| * scale); L2 keeps |c|^2 beside the uniform scale
            tc::mul_rows_kernel<<<static_cast<uint32_t>((static_cast<uint64_t>(st->n_pad) + 127) / 128), 128, 0, s>>>(st->d_aux, st->d_aux2, st->n_pad);
    } else {
        tc::split_tf32_kernel<<<tc_blocks_for(static_cast<uint64_t>(st->n_pad) * kp), 256, 0, s>>>(ix->d_centroids, ix->cent_ld, ix->dim, ix->nlist, st->n_pad, kp,
                                                                                                static_cast<float*>(st->d_x));
    }
    ANNB_CUDA_CHECK(cudaGetLastError());
    ANNB_TRY(tc_make_tmap(&st->tm_x, st->d_x, 2ull * st->n_pad, kp, elem));
    ix->device_bytes += st->bytes;
    return ANNB_OK;
}

void tc_coarse_destroy(annb_index* ix) {
    if (!ix->tc_coarse) return;
    ix->device_bytes -= std::min<uint64_t>(ix->device_bytes, ix->tc_coarse->bytes);
    cudaFree(ix->tc_coarse->d_x);
    cudaFree(ix->tc_coarse->d_aux);
    cudaFree(ix->tc_coarse->d_aux2);
    ix->tc_coarse->q_scale.release();
    ix->tc_coarse->q_op.release();
    ix->tc_coarse->dense.release();
    ix->tc_coarse->dense_gm.release();
    delete ix->tc_coarse;
    ix->tc_coarse = nullptr;
}

bool tc_coarse_supported(const annb_index* ix) { return ix->tc_coarse != nullptr; }

template <int MET>
static int launch_coarse_select(const tc::CoarseSelectParams& c, cudaStream_t s) {
    auto kern = c.gmin ? tc::coarse_select_gm_kernel<MET> : tc::coarse_select_kernel<MET>;
    const uint32_t warps = (c.staged_words && !c.gmin) ? 4 : 8;
    const size_t per_warp = c.gmin ? tc::coarse_gm_warp_bytes(c.cmax, c.dense_ld >> 3, c.c_ld, c.dim) : tc::coarse_select_warp_bytes(c.cmax, c.staged_words);
    const size_t smem = per_warp * warps;
    ANNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<static_cast<uint32_t>((c.nq + warps - 1) / warps), warps * 32, smem, s>>>(c);
    ANNB_CUDA_CHECK(cudaGetLastError());
    return ANNB_OK;
}

// Ranked prefix of the centroid table for every query: d_ranked[nq][pitch] ascending (distance, cell) keys, sentinels
// where a rank could not be certified.
int tc_coarse_rank(annb_index* ix, const float* d_route, uint32_t route_ld, uint64_t nq, uint32_t pitch, uint64_t* d_ranked, cudaStream_t s, const CoarseWalk* walk) {
    TcState* st = ix->tc_coarse;
    const uint32_t kp = st->kp_elems;
    const uint32_t nq_pad = static_cast<uint32_t>(round_up<uint64_t>(nq, tc::BM));
    const bool f16 = st->kind == tc::KIND_F16X3;
    const uint32_t elem = f16 ? 2u : 4u;
    ANNB_TRY(st->q_op.ensure(2ull * nq_pad * kp * elem));
    if (f16) {
        ANNB_TRY(st->q_scale.ensure(static_cast<uint64_t>(nq_pad) * 4));
        tc::split_f16_kernel<<<tc_blocks_for(static_cast<uint64_t>(nq_pad) * 32), 256, 0, s>>>(d_route, route_ld, ix->dim, nq, nq_pad, kp, st->q_op.as<__half>(),
                                                                                            st->q_scale.as<float>());
    } else {
        tc::split_tf32_kernel<<<tc_blocks_for(static_cast<uint64_t>(nq_pad) * kp), 256, 0, s>>>(d_route, route_ld, ix->dim, nq, nq_pad, kp, st->q_op.as<float>());
    }
    ANNB_CUDA_CHECK(cudaGetLastError());
    CUtensorMap tmq;
    ANNB_TRY(tc_make_tmap(&tmq, st->q_op.p, 2ull * nq_pad, kp, elem));
    const uint64_t q_tiles = nq_pad / tc::BM, db_tiles = st->n_pad / tc::BN;
    // One CTA per SM: whole waves of CTAs against the per-CTA fixed cost (about two tiles' worth of query staging and pipeline fill)
    // -- 79 query tiles: 5 splits = 2.7 waves of 7 tiles instead of 4 splits = 2.1 waves (three in practice) of 8
    const uint32_t splits_req = pick_splits(q_tiles, db_tiles, 0, 2);
    const uint64_t tiles_per = (db_tiles + splits_req - 1) / splits_req;
    const uint32_t splits = static_cast<uint32_t>((db_tiles + tiles_per - 1) / tiles_per);
    const size_t fixed = 256 + 2 * 8 * 64 * 4, budget = 227 * 1024;
    const uint32_t stages = static_cast<uint32_t>(std::min<size_t>(8, (budget - fixed) / (2 * tc::SLAB_TILE)));
    const size_t smem = static_cast<size_t>(stages) * 2 * tc::SLAB_TILE + fixed;
    ANNB_TRY(st->dense.ensure(nq * static_cast<uint64_t>(st->n_pad) * 4));
    tc::Params p{};
    p.nq = nq; p.n_rows = ix->nlist; p.nq_pad = nq_pad; p.n_pad = st->n_pad; p.nslab = st->nslab; p.n_stages = stages;
    p.n_splits = splits; p.rows_per_split = tiles_per * tc::BN; p.a_pieces = 2; p.aux = st->d_aux;
    p.q_op = st->q_op.as<void>(); p.kp = kp; p.dense = st->dense.as<float>(); p.dense_ld = st->n_pad;
    p.aux2 = st->d_aux2; p.q_inv_scale = st->q_scale.as<float>(); p.db_inv_scale = st->db_inv_scale;
    // Select from group minima (coarse_select_gm_kernel) when the selected groups are expected to hold well under cmax cells at
    // or below the threshold: G groups carry about G (1 + 7 G / nlist) of them.
    const uint32_t cmax = next_pow2(pitch + 1);
    uint32_t glimit = 0;
    if (ix->opt_ivf_coarse_gm && pitch < ix->nlist && st->n_pad / 8 <= 2048) {
        auto expect = [&](uint32_t g) { return g * (1.0 + 7.0 * g / ix->nlist); };
        if (expect(pitch) * 1.15 <= cmax) {
            glimit = pitch;
            while (glimit + 1 <= cmax && expect(glimit + 1) <= 0.9 * cmax) glimit++;
        }
    }
    if (glimit) {
        ANNB_TRY(st->dense_gm.ensure(nq * static_cast<uint64_t>(st->n_pad / 8) * 4));
        p.dense_gm = st->dense_gm.as<float>();
        // the select then reads single pieces: store the matrix in the blocked layout whose epilogue stores coalesce (whole query tiles)
        ANNB_TRY(st->dense.ensure(static_cast<uint64_t>(nq_pad) * st->n_pad * 4));
        p.dense = st->dense.as<float>();
        p.dense_blocked = ix->opt_ivf_coarse_blocked ? 1u : 0u;
    }
    {
        dim3 grid(static_cast<uint32_t>(q_tiles), splits);
        if (f16 && ix->metric == ANNB_L2) {
            auto kern = tc::flat_tc_kernel<tc::KIND_F16X3, 16, MET_L2, true, true>;
            ANNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            kern<<<grid, tc::NUM_THREADS, smem, s>>>(tmq, st->tm_x, p);
        } else if (f16) {
            auto kern = tc::flat_tc_kernel<tc::KIND_F16X3, 16, MET_COS, true, true>;
            ANNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            kern<<<grid, tc::NUM_THREADS, smem, s>>>(tmq, st->tm_x, p);
        } else if (ix->metric == ANNB_L2) {
            auto kern = tc::flat_tc_kernel<tc::KIND_TF32X3, 16, MET_L2, true, true>;
            ANNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            kern<<<grid, tc::NUM_THREADS, smem, s>>>(tmq, st->tm_x, p);
        } else {
            auto kern = tc::flat_tc_kernel<tc::KIND_TF32X3, 16, MET_COS, true, true>;
            ANNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            kern<<<grid, tc::NUM_THREADS, smem, s>>>(tmq, st->tm_x, p);
        }
        ANNB_CUDA_CHECK(cudaGetLastError());
    }
    tc::CoarseSelectParams c{};
    c.dense = st->dense.as<float>(); c.dense_ld = st->n_pad; c.nq = nq; c.nlist = ix->nlist; c.pitch = pitch; c.cmax = cmax;
    // Staging the row of values in shared memory (one global pass instead of ~5 L2 passes) was measured SLOWER at nlist 4096,
    // 10k queries: 0.41 ms against 0.27 ms per launch -- 16 KB per warp leaves 12 resident warps per SM where the L2 variant
    // keeps 64, and the select is a chain of dependent passes that only occupancy hides.  Kept behind option ivf_coarse_stage.
    c.staged_words = (ix->opt_ivf_coarse_stage && ix->nlist <= 8192) ? round_up(ix->nlist, 4u) : 0u;
    if (glimit) { c.gmin = st->dense_gm.as<float>(); c.glimit = glimit; c.staged_words = st->n_pad / 8; c.blocked = p.dense_blocked; }
    c.queries = d_route; c.q_ld = route_ld; c.centroids = ix->d_centroids; c.c_ld = ix->cent_ld; c.centroid_norms = ix->d_centroid_norms;
    c.dim = ix->dim; c.eps = tc_cert_eps(ix, st->kind, kp, 2, false); c.cnorm_max = ix->tc_cnorm_max; c.ranked = d_ranked;
    if (walk != nullptr) { c.do_walk = 1; c.walk = *walk; }
    if (ix->metric == ANNB_L2) ANNB_TRY(launch_coarse_select<MET_L2>(c, s));
    else if (ix->dtype == ANNB_SQ8) ANNB_TRY(launch_coarse_select<MET_COS_PRENORM>(c, s));
    else ANNB_TRY(launch_coarse_select<MET_COS>(c, s));
    ix->stat_launches += 3;
    return ANNB_OK;
}

// ----------------------------------------------------------------------------------------------- build-side assignment (host)
struct TcAssignState {
    uint32_t dim = 0, nlist = 0, kp = 0, nslab = 0, n_pad = 0;
    void* d_x = nullptr;          // centroid table as stacked tf32 hi / lo operand
    float* d_aux = nullptr;       // [n_pad + BN]
    uint32_t* d_cmax = nullptr;   // bit pattern of max |c|^2
    CUtensorMap tm_x;
    DevBuf q_op, part, gtau;
};

// nullptr state (and ANNB_OK) when the shape is not covered: small tables are cheap on the CUDA cores anyway.
int tc_assign_create(TcAssignState** out, uint32_t dim, uint32_t nlist) {
    *out = nullptr;
    const uint32_t kp = round_up(dim, tc::SLAB_BYTES / 4u);
    if (nlist < 512 || kp * 4 > 512) return ANNB_OK;
    TcAssignState* st = new TcAssignState();
    st->dim = dim; st->nlist = nlist; st->kp = kp; st->nslab = kp / (tc::SLAB_BYTES / 4u);
    st->n_pad = round_up<uint32_t>(nlist, tc::BN);
    const uint64_t aux_rows = static_cast<uint64_t>(st->n_pad) + tc::BN;
    cudaError_t e = cudaMalloc(&st->d_x, 2ull * st->n_pad * kp * 4);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&st->d_aux), aux_rows * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&st->d_cmax), 4);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_last_error(std::string("cudaMalloc assignment operand: ") + cudaGetErrorString(e));
        tc_assign_destroy(st);
        return ANNB_ERR_OUT_OF_MEMORY;
    }
    int rc = tc_make_tmap(&st->tm_x, st->d_x, 2ull * st->n_pad, kp, 4);
    if (rc != ANNB_OK) { tc_assign_destroy(st); return rc; }
    *out = st;
    return ANNB_OK;
}

void tc_assign_destroy(TcAssignState* st) {
    if (!st) return;
    cudaFree(st->d_x);
    cudaFree(st->d_aux);
    cudaFree(st->d_cmax);
    st->q_op.release();
    st->part.release();
    st->gtau.release();
    delete st;
}

// (Re)loads the centroid table: Lloyd iterations call this once per iteration.
int tc_assign_set_centroids(TcAssignState* st, const float* d_c, uint32_t c_ld, const float* d_assign_aux, bool cosine, cudaStream_t s) {
    const uint64_t aux_rows = static_cast<uint64_t>(st->n_pad) + tc::BN;
    tc::split_tf32_kernel<<<tc_blocks_for(static_cast<uint64_t>(st->n_pad) * st->kp), 256, 0, s>>>(d_c, c_ld, st->dim, st->nlist, st->n_pad, st->kp, static_cast<float*>(st->d_x));
    tc::assign_tc_aux_kernel<<<static_cast<uint32_t>((aux_rows + 127) / 128), 128, 0, s>>>(d_assign_aux, st->nlist, aux_rows, cosine ? 1 : 0, st->d_aux);
    ANNB_CUDA_CHECK(cudaMemsetAsync(st->d_cmax, 0, 4, s));
    if (!cosine) tc::aux_max_kernel<<<32, 256, 0, s>>>(d_assign_aux, st->nlist, st->d_cmax);
    ANNB_CUDA_CHECK(cudaGetLastError());
    return ANNB_OK;
}

// Assigns rows [0, nr) of d_x; d_uncert[0] receives the number of rows that must be redone exactly, d_uncert[1..] their numbers.
int tc_assign_run(TcAssignState* st, const float* d_x, uint32_t x_ld, uint64_t nr, const float* d_c, uint32_t c_ld, const float* d_assign_aux,
                  bool cosine, uint32_t* d_assign, uint32_t* d_uncert, cudaStream_t s) {
    constexpr uint32_t AKP = 16;
    const uint32_t kp = st->kp;
    const uint32_t nq_pad = static_cast<uint32_t>(round_up<uint64_t>(nr, tc::BM));
    ANNB_TRY(st->q_op.ensure(2ull * nq_pad * kp * 4));
    tc::split_tf32_kernel<<<tc_blocks_for(static_cast<uint64_t>(nq_pad) * kp), 256, 0, s>>>(d_x, x_ld, st->dim, nr, nq_pad, kp, st->q_op.as<float>());
    ANNB_CUDA_CHECK(cudaGetLastError());
    CUtensorMap tmq;
    ANNB_TRY(tc_make_tmap(&tmq, st->q_op.p, 2ull * nq_pad, kp, 4));
    const uint64_t q_tiles = nq_pad / tc::BM, db_tiles = st->n_pad / tc::BN;
    const uint32_t splits_req = pick_splits(q_tiles, db_tiles, 0);
    const uint64_t tiles_per = (db_tiles + splits_req - 1) / splits_req;
    const uint32_t splits = static_cast<uint32_t>((db_tiles + tiles_per - 1) / tiles_per);
    const size_t fixed = 256 + 2 * 8 * 64 * 4, budget = 227 * 1024;
    const uint32_t stages = static_cast<uint32_t>(std::min<size_t>(8, (budget - fixed) / (2 * tc::SLAB_TILE)));
    const size_t smem = static_cast<size_t>(stages) * 2 * tc::SLAB_TILE + fixed;
    ANNB_TRY(st->part.ensure(nr * 2ull * splits * AKP * 8));
    ANNB_TRY(st->gtau.ensure(static_cast<uint64_t>(nq_pad) * 4));
    ANNB_CUDA_CHECK(cudaMemsetAsync(st->gtau.p, 0xFF, static_cast<uint64_t>(nq_pad) * 4, s));
    ANNB_CUDA_CHECK(cudaMemsetAsync(d_uncert, 0, 4, s));
    tc::Params p{};
    p.nq = nr; p.n_rows = st->nlist; p.nq_pad = nq_pad; p.n_pad = st->n_pad; p.nslab = st->nslab; p.n_stages = stages;
    p.n_splits = splits; p.rows_per_split = tiles_per * tc::BN; p.a_pieces = 2; p.aux = st->d_aux;
    p.q_op = st->q_op.as<void>(); p.kp = kp; p.part_keys = st->part.as<uint64_t>(); p.gtau = st->gtau.as<uint32_t>();
    dim3 grid(static_cast<uint32_t>(q_tiles), splits);
    if (cosine) ANNB_TRY((launch_tc<tc::KIND_TF32X3, AKP, MET_COS, true>(tmq, st->tm_x, p, grid, smem, s)));
    else ANNB_TRY((launch_tc<tc::KIND_TF32X3, AKP, MET_L2, true>(tmq, st->tm_x, p, grid, smem, s)));
    tc::AssignSelectParams a{};
    a.part_keys = st->part.as<uint64_t>(); a.parts = 2 * splits; a.kp = AKP; a.gtau = st->gtau.as<uint32_t>(); a.n = nr;
    a.rows = d_x; a.x_ld = x_ld; a.centroids = d_c; a.c_ld = c_ld; a.aux = d_assign_aux; a.dim = st->dim; a.cosine = cosine ? 1 : 0;
    // error budget of the selection value against the reference-order score: 3xTF32 products and the tensor core's
    // accumulation (< 8.5 * 2^-20 |x||c|) plus the f32 rounding of the reference's own sequential sum
    // (<= dim * 2^-24 |x||c| in the worst case); 2^-16 (|x| + |c|max)^2 >= 2^-14 |x||c| covers both up to dim 512
    a.eps = 1.52587890625e-05f;
    a.cmax_sq_bits = st->d_cmax; a.assign_out = d_assign; a.uncert = d_uncert;
    tc::assign_select_kernel<<<static_cast<uint32_t>((nr + 7) / 8), 256, 0, s>>>(a);
    ANNB_CUDA_CHECK(cudaGetLastError());
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
