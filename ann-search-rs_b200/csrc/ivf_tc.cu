// ivf_tc.cu -- IVF list scan on the tensor cores ("grouped" distance tiles).
//
// The (query, probe-rank) pairs of a batch are bucketed by inverted list on the device (api.cu).  One task = one list x a
// group of up to 128 of the queries that probe it.  Persistent CTAs pull tasks from an atomic counter and run the same
// warp-specialised pipeline as flat_tc_kernel, with three differences:
//   * the 128 query rows of a task are *gathered* by the epilogue threads (thread = TMEM lane = one (query, rank) pair)
//     from the batch's pre-split operand pieces (tf32 hi/lo, three bf16 terms, or the int8 codes) into TMEM (TS-mode MMA);
//   * the database operand is the list's own slab [offsets[c], offsets[c+1]) of the index rows themselves, walked in
//     128-row tiles by TMA (f32: two transform warps derive the lo operand from every landed slab, the raw slab is hi);
//     rows of the last tile that belong to the next list are masked with a NaN row constant;
//   * each (query, rank) keeps its own k' candidates; ivf rerank recomputes the merged survivors exactly in the reference's
//     arithmetic and orders them by (distance, list position) like SortedBuffer (src/cpu/ivf.rs:367-381).
// Tasks arrive longest list first (ivf_pair_offsets_kernel with the index's d_list_order), and an epilogue warp hands its
// accumulator stage back as soon as the tile's values are in registers, before any select work.
// Replaces compute_ivf_mega_* + radix_select_ivf_topk of the reference (src/gpu/dist_gpu.rs:922-1355, src/gpu/topk_gpu.rs:599-832).
#include <cuda.h>

#include <algorithm>
#include <cstring>

#include "flat_tc.hpp"
#include "index.hpp"
#include "tc_common.cuh"

namespace annb {

namespace tc {

struct IvfTcParams {
    const uint8_t* queries;   // prepared scan queries: f32 rows / int8 codes (q_bytes pitch)
    uint32_t q_bytes;
    const void* q_op;         // f32 / bf16 lists: the batch's queries pre-split into operand pieces, [pieces][nq][kp] (tf32 hi, lo / bf16 q0, q1, q2)
    uint64_t nq;
    uint32_t dim;
    uint32_t nslab, n_stages;
    uint32_t n_pad;           // rows per piece of the stacked database operand
    const float* aux;         // per stored row (shard-local list order): L2 |x|^2, cosine -1/|x|
    const uint64_t* offsets;  // global CSR offsets
    uint64_t shard_row0;
    uint32_t nlist;
    const uint32_t* pair_off;
    const uint32_t* task_off;  // prefix sum of ceil(pairs per list / 128)
    const uint2* pairs;        // (query, rank) grouped by list
    const uint4* tasks;        // [2 * n_tasks] task records written by ivf_pair_offsets_kernel
    uint32_t* task_counter;
    uint32_t prefetch_task;    // 1: the producer lane claims / fetches the next task during the last tiles of the current one
    uint32_t probe_pitch;
    uint32_t bf16_terms;       // bf16 lists: bf16 terms of the f32 query (2 or 3)
    uint64_t* part_keys;       // [nq][probe_pitch][2][KP]
    uint32_t* gtau;            // [nq] shared pruning threshold
    const float* aux2;         // KIND_F16X3: per stored row 1 / (its power-of-two operand scale)
    const float* q_inv_scale;  // KIND_F16X3: per query 1 / (its power-of-two operand scale), [nq]
    float db_inv_scale;        // KIND_F16X3, L2: 1 / (the list operand's uniform power-of-two scale)
    uint32_t lo_smem;          // f32 lists with rows of 129 .. 256 elements: only the hi query piece fits TMEM beside two accumulator stages;
                               // the lo piece is gathered into shared memory (swizzled K slabs) and its term is an SS-mode MMA
    unsigned long long* dbg;   // optional [8]: CTA 0 cycle counters {total, schedule, gather, epi wait-tfull, mma wait-queries, mma wait-data, mma wait-tempty, tasks << 32 | tiles}
};

// f32 lists: the kernel streams the index's own f32 rows (one 128 B K-slab per TMA load) and two extra "transform" warps
// split every landed slab into tf32 hi (in place) and lo (second slab of the stage) before the MMA reads it -- the list is
// read from HBM once per task at 4 B per element instead of 8 B from a pre-split copy, and no such copy is stored.
// Warp roles: 0 TMA, 1 MMA, [2, 3 transform,] then the 8 epilogue warps.
constexpr int XF_THREADS = 64;
template <int KIND> constexpr int ivf_tc_threads() { return KIND == KIND_TF32X3 ? NUM_THREADS + XF_THREADS : NUM_THREADS; }

template <int KIND, int KP, int MET>
__global__ void __launch_bounds__(ivf_tc_threads<KIND>(), 1) ivf_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const IvfTcParams p) {
    // KIND_F16X3 (f32 lists, 3xFP16): the operand is a pre-split copy of the lists, stacked fp16 hi / lo pieces of the rows scaled by
    // their powers of two (split_f16_kernel; 4 B per element like the f32 rows themselves) -- a ring stage is one hi and one lo K slab
    // (64 elements of K), no transform warps.  (Converting the raw f32 slabs inside the kernel was tried first: two transform warps
    // need ~2 000 cycles per tile for it and paced the kernel at 3.2 ms where 3xTF32 takes 2.1.)
    constexpr bool XFORM = (KIND == KIND_TF32X3);
    constexpr bool F16 = (KIND == KIND_F16X3);
    constexpr int NB = (XFORM || F16) ? 2 : 1;
    constexpr int ELEM = XFORM ? 4 : (KIND == KIND_I8 ? 1 : 2);      // element size of what TMA loads
    constexpr int SLAB_ELEMS = SLAB_BYTES / ELEM;
    constexpr int KSTEPS = 4;
    // TMEM: query pieces at column 0 (f32 hi / lo 2 x 128 columns, bf16 terms 64 or 128 each, int8 codes <= 128), accumulator
    // ring behind them: three stages when the pieces fit 128 columns, two otherwise
    const bool lo_s = KIND == KIND_TF32X3 && p.lo_smem != 0;
    const uint32_t PIECE_COLS = (KIND == KIND_TF32X3) ? (lo_s ? p.nslab * 32u : 128u)
                                : (F16 ? p.nslab * 32u : (KIND == KIND_I8 ? 128u : (p.nslab > 2u ? 128u : 64u)));
    const uint32_t q_cols = (KIND == KIND_TF32X3) ? (lo_s ? PIECE_COLS : 256u) : (F16 ? 2u * PIECE_COLS : (KIND == KIND_I8 ? 128u : p.bf16_terms * PIECE_COLS));
    const uint32_t NACC = q_cols <= 128u ? 3u : 2u;
    const uint32_t ACC_COL0 = q_cols <= 128u ? 128u : 256u;
    constexpr uint32_t idesc = make_idesc(KIND);
    constexpr uint32_t SLAB_DESC = SLAB_TILE >> 4;

    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* s_qlo = smem;                                                  // lo_smem: [nslab] slabs holding the task's lo query piece
    uint8_t* s_x = lo_s ? smem + static_cast<size_t>(p.nslab) * SLAB_TILE : smem;   // [n_stages][NB] slabs
    uint8_t* s_tail = s_x + static_cast<size_t>(p.n_stages) * NB * SLAB_TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_tail);
    uint64_t* bar_full = bars;                     // [n_stages]
    uint64_t* bar_empty = bars + p.n_stages;       // [n_stages]
    uint64_t* bar_xf = bars + 2 * p.n_stages;      // [n_stages] slab split into hi / lo (f32 lists)
    uint64_t* bar_q = bars + 3 * p.n_stages;       // [1]  queries of the current task are in TMEM
    uint64_t* bar_tfull = bar_q + 1;               // [3]
    uint64_t* bar_tempty = bar_tfull + 3;          // [3]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_tempty + 3);
    uint32_t* s_task = s_tmem + 1;                 // [4]: list, pair0, n_in_group, valid flag
    uint64_t* s_rows = reinterpret_cast<uint64_t*>(s_tmem + 6);  // [2]: r_begin, r_end (8-byte aligned: bars + ... even count)
    float* s_aux_all = reinterpret_cast<float*>(s_tail + 512);   // [2 banks][8 epilogue warps][64] row constants of the warp's current half tile

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.n_stages; s++) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); mbar_init(bar_xf + s, XF_THREADS); }
        mbar_init(bar_q, EPI_THREADS);
        for (uint32_t a = 0; a < NACC; a++) { mbar_init(bar_tfull + a, 1); mbar_init(bar_tempty + a, EPI_THREADS); }
        fence_barrier_init();
        fence_proxy_async();
        tma_prefetch_desc(&tm_x);
    }
    if (warp == 1) tmem_alloc(s_tmem, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const uint32_t total_tasks = p.task_off[p.nlist];

    // role-local running counters: the smem ring and the accumulator ring keep rolling across tasks
    // ring position / accumulator stage and their phase bits: every role advances its own copies identically, slab by slab and tile
    // by tile (incrementally: no division on the per-slab / per-tile chains)
    uint32_t r_stage = 0, r_ph = 0, r_acc = 0, r_aph = 0;
    auto next_slab = [&]() { if (++r_stage == p.n_stages) { r_stage = 0; r_ph ^= 1u; } };
    auto next_tile = [&]() { if (++r_acc == NACC) { r_acc = 0; r_aph ^= 1u; } };
    uint32_t task_no = 0;   // tasks done by this CTA (parity of bar_q)

    // epilogue-only state
    const uint32_t quarter = warp & 3u;
    const uint32_t row_in_tile = quarter * 32 + lane;
    constexpr uint32_t EPI_WARP0 = XFORM ? 2 + XF_THREADS / 32 : 2;
    const uint32_t half = (warp >= EPI_WARP0) ? ((warp - EPI_WARP0) >> 2) : 0;
    TopList<KP> top;
    float scratch[64];

    const bool dbg_on = TC_COUNTERS && p.dbg != nullptr && blockIdx.x == 0;
    long long c_sched = 0, c_gather = 0, c_tfull = 0, c_wq = 0, c_wdata = 0, c_wtempty = 0, c_tasks = 0, c_tiles = 0;
    const long long c_start = tc_clock();
    // Thread 0 (the TMA producer's issuing lane) claims and fetches the NEXT task while it streams the last tiles of the current one:
    // the claim (an atomic on the shared counter) a few tiles before the end, the record two tiles before the end -- under the scan's
    // own memory traffic each of the two dependent round trips takes microseconds, which used to sit between two CTA barriers with
    // every role idle (19 % of the kernel).  Claiming a handful of tiles early, not a whole task early, keeps the tail of the dynamic
    // schedule as short as before.
    // The record travels by cp.async into a spare shared-memory slot: no registers are held for it across the task.
    uint32_t nx_task = 0, nx_state = 0;            // 0: nothing in flight, 1: next task claimed, 2: ... and its record requested
    uint4* s_next = reinterpret_cast<uint4*>(s_tail + 448);   // [2]
    for (;;) {
        __syncthreads();   // every role has finished the previous task (TMEM query region and s_task are reusable)
        const long long c_t0 = tc_clock();
        if (threadIdx.x == 0) {
            if (nx_state == 0) nx_task = atomicAdd(p.task_counter, 1u);
            const uint32_t task = nx_task;
            if (task < total_tasks) {
                uint4 a, b;
                if (nx_state == 2) { cp_async_wait<0>(); a = s_next[0]; b = s_next[1]; }
                else { a = __ldg(p.tasks + 2 * static_cast<size_t>(task)); b = __ldg(p.tasks + 2 * static_cast<size_t>(task) + 1); }
                nx_state = 0;
                s_task[0] = a.x;
                s_task[1] = a.y;
                s_task[2] = a.z;
                s_task[3] = 1;
                s_rows[0] = static_cast<uint64_t>(b.x) | (static_cast<uint64_t>(b.y) << 32);
                s_rows[1] = static_cast<uint64_t>(b.z) | (static_cast<uint64_t>(b.w) << 32);
            } else {
                s_task[3] = 0;
            }
        }
        __syncthreads();
        if (s_task[3] == 0) break;
        const long long c_t1 = tc_clock();
        c_sched += c_t1 - c_t0;
        const uint32_t pair0 = s_task[1], n_in_group = s_task[2];
        const uint64_t r_begin = s_rows[0], r_end = s_rows[1];
        const uint32_t n_tiles = static_cast<uint32_t>((r_end - r_begin + BN - 1) / BN);

        if (warp == 0) {
            // ================================================================= TMA producer
            if (lane == 0) {
                const uint32_t claim_at = (p.prefetch_task && n_tiles > 6) ? n_tiles - 6 : 0u, fetch_at = n_tiles > 2 ? n_tiles - 2 : n_tiles - 1;
                for (uint32_t t = 0; t < n_tiles; t++) {
                    if (p.prefetch_task) {
                        if (t == claim_at) { nx_task = atomicAdd(p.task_counter, 1u); nx_state = 1; }
                        if (t == fetch_at && nx_state == 1 && nx_task < total_tasks) {
                            cp_async16(s_next, p.tasks + 2 * static_cast<size_t>(nx_task));
                            cp_async16(s_next + 1, p.tasks + 2 * static_cast<size_t>(nx_task) + 1);
                            cp_async_commit();
                            nx_state = 2;
                        }
                    }
                    const uint32_t row0 = static_cast<uint32_t>(r_begin) + t * BN;
                    for (uint32_t s = 0; s < p.nslab; s++, next_slab()) {
                        const uint32_t stage = r_stage, ph = r_ph;
                        mbar_wait(bar_empty + stage, ph ^ 1u);
                        if (XFORM) {   // raw f32 slab; rows past the shard's end are zero filled by TMA
                            mbar_expect_tx(bar_full + stage, SLAB_TILE);
                            tma_load_2d(smem_u32(s_x + static_cast<size_t>(stage) * NB * SLAB_TILE), &tm_x, bar_full + stage, s * SLAB_ELEMS, row0);
                        } else {
                            mbar_expect_tx(bar_full + stage, NB * SLAB_TILE);
                            for (int b = 0; b < NB; b++)
                                tma_load_2d(smem_u32(s_x + (static_cast<size_t>(stage) * NB + b) * SLAB_TILE), &tm_x, bar_full + stage, s * SLAB_ELEMS,
                                            b * p.n_pad + row0);
                        }
                    }
                }
            }                              // (the other lanes of the producer warp never use their ring position)
            __syncwarp();
        } else if (warp == 1) {
            // ================================================================= MMA issuer (converged warp, elected issue)
            mbar_wait_timed(bar_q, task_no & 1u, c_wq);
            tc_fence_after();
            const uint32_t x_desc0 = make_smem_desc(smem_u32(s_x));   // low descriptor word
            const uint32_t qlo_desc0 = make_smem_desc(smem_u32(s_qlo));
            for (uint32_t t = 0; t < n_tiles; t++, next_tile()) {
                const uint32_t acc = r_acc, aph = r_aph;
                mbar_wait_timed(bar_tempty + acc, aph ^ 1u, c_wtempty);
                tc_fence_after();
                const uint32_t tmem_c = tmem_base + ACC_COL0 + acc * BN;
                for (uint32_t s = 0; s < p.nslab; s++, next_slab()) {
                    const uint32_t stage = r_stage, ph = r_ph;
                    mbar_wait_timed((XFORM ? bar_xf : bar_full) + stage, ph, c_wdata);
                    tc_fence_after();
                    const uint32_t xd = x_desc0 + stage * NB * SLAB_DESC;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < KSTEPS; k++) {
                            const uint32_t first = (s | static_cast<uint32_t>(k)) != 0 ? 1u : 0u;
                            const uint32_t a0 = tmem_base + s * 32 + k * 8;
                            if (F16) {
                                umma_ts<KIND>(tmem_c, a0, xd + 2 * k, idesc, first);
                                umma_ts<KIND>(tmem_c, a0 + PIECE_COLS, xd + 2 * k, idesc, 1u);
                                umma_ts<KIND>(tmem_c, a0, xd + SLAB_DESC + 2 * k, idesc, 1u);
                            } else if (KIND == KIND_TF32X3) {
                                umma_ts<KIND>(tmem_c, a0, xd + 2 * k, idesc, first);
                                if (lo_s) umma<KIND>(tmem_c, qlo_desc0 + s * SLAB_DESC + 2 * k, xd + 2 * k, idesc, 1u);   // Qlo from shared memory
                                else umma_ts<KIND>(tmem_c, a0 + PIECE_COLS, xd + 2 * k, idesc, 1u);
                                umma_ts<KIND>(tmem_c, a0, xd + SLAB_DESC + 2 * k, idesc, 1u);
                            } else if (KIND == KIND_I8) {
                                umma_ts<KIND>(tmem_c, a0, xd + 2 * k, idesc, first);
                            } else {
                                umma_ts<KIND>(tmem_c, a0, xd + 2 * k, idesc, first);
                                umma_ts<KIND>(tmem_c, a0 + PIECE_COLS, xd + 2 * k, idesc, 1u);
                                if (p.bf16_terms > 2) umma_ts<KIND>(tmem_c, a0 + 2 * PIECE_COLS, xd + 2 * k, idesc, 1u);
                            }
                        }
                        umma_commit(bar_empty + stage);
                        if (s + 1 == p.nslab) umma_commit(bar_tfull + acc);
                    }
                    __syncwarp();
                }
            }
        } else if (XFORM && warp < EPI_WARP0) {
            // ================================================================= transform (2 warps): slab -> tf32 lo tile (hi = the raw slab)
            const uint32_t xt = threadIdx.x - 2 * 32;
            for (uint32_t t = 0; t < n_tiles; t++) {
                for (uint32_t s = 0; s < p.nslab; s++, next_slab()) {
                    const uint32_t stage = r_stage, ph = r_ph;
                    mbar_wait(bar_full + stage, ph);
                    uint8_t* raw = s_x + static_cast<size_t>(stage) * NB * SLAB_TILE;
                    // elementwise, so the slab's swizzled layout carries over: thread i owns 16-byte chunks i, i + 64, ...
#pragma unroll
                    for (int j = 0; j < SLAB_TILE / 16 / XF_THREADS; j++) {
                        const uint32_t off = (xt + XF_THREADS * j) * 16;
                        // hi needs no work: the tensor core reads only the upper 19 bits of a tf32 operand, i.e. it truncates x in
                        // place.  lo = x - trunc(x) is exact in f32 (<= 13 significant bits); adding half a tf32 ulp to its bit
                        // pattern makes the hardware's truncation round it to nearest.  |lo| < 2^-10 |x|, error <= 2^-21 |x|.
                        const uint4 x = *reinterpret_cast<const uint4*>(raw + off);
                        uint4 l;
                        l.x = __float_as_uint(__fsub_rn(__uint_as_float(x.x), __uint_as_float(x.x & 0xFFFFE000u))) + 0x1000u;
                        l.y = __float_as_uint(__fsub_rn(__uint_as_float(x.y), __uint_as_float(x.y & 0xFFFFE000u))) + 0x1000u;
                        l.z = __float_as_uint(__fsub_rn(__uint_as_float(x.z), __uint_as_float(x.z & 0xFFFFE000u))) + 0x1000u;
                        l.w = __float_as_uint(__fsub_rn(__uint_as_float(x.w), __uint_as_float(x.w & 0xFFFFE000u))) + 0x1000u;
                        *reinterpret_cast<uint4*>(raw + SLAB_TILE + off) = l;
                    }
                    fence_proxy_async();           // generic-proxy writes -> visible to the tensor core's async-proxy reads
                    mbar_arrive(bar_xf + stage);
                }
            }
        } else {
            // ================================================================= epilogue (8 warps)
            const uint2 pr = (row_in_tile < n_in_group) ? p.pairs[pair0 + row_in_tile] : make_uint2(0xFFFFFFFFu, 0u);
            const bool has_query = pr.x != 0xFFFFFFFFu;
            const bool warp_has_query = __any_sync(0xFFFFFFFFu, has_query);   // groups are filled from row 0: whole warps may be idle
            // ---- gather this lane's query, split it, store it into TMEM ----
            {
                const float* qrow = reinterpret_cast<const float*>(p.queries + static_cast<uint64_t>(has_query ? pr.x : 0) * p.q_bytes);
                const uint32_t kp = p.nslab * SLAB_ELEMS;
                if (KIND == KIND_I8) {
                    // the scan queries are already int8 codes (zero padded to 16): four codes per TMEM column, written by half 0
                    if (half == 0) {
                        const uint32_t* qw = reinterpret_cast<const uint32_t*>(qrow);
                        const uint32_t tq = tmem_base + ((quarter * 32u) << 16);
                        for (uint32_t c = 0; c < kp / 4; c += 32) {
                            uint32_t w[32];
#pragma unroll
                            for (int j = 0; j < 32; j++) w[j] = (has_query && 4 * (c + j) < p.q_bytes) ? __ldg(qw + c + j) : 0u;
                            tmem_st32(tq + c, w);
                        }
                    }
                } else if (KIND == KIND_TF32X3 && lo_s && half == 1) {
                    // wide rows: half 1 writes the lo piece into the shared-memory K slabs in the layout the MMA descriptor reads
                    // (SWIZZLE_128B: row r of a slab at r * 128 B, its 16-byte chunk c at position c ^ (r & 7))
                    const uint4* src = reinterpret_cast<const uint4*>(static_cast<const float*>(p.q_op) + (static_cast<uint64_t>(p.nq) + (has_query ? pr.x : 0)) * kp);
                    for (uint32_t sl = 0; sl < p.nslab; sl++) {
                        uint4 x[8];
#pragma unroll
                        for (int c = 0; c < 8; c++) x[c] = has_query ? __ldg(src + sl * 8 + c) : make_uint4(0u, 0u, 0u, 0u);
                        uint8_t* dst = s_qlo + static_cast<size_t>(sl) * SLAB_TILE + row_in_tile * 128u;
#pragma unroll
                        for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(dst + ((static_cast<uint32_t>(c) ^ (row_in_tile & 7u)) << 4)) = x[c];
                    }
                    fence_proxy_async();           // generic-proxy writes -> visible to the tensor core's async-proxy reads
                } else if (XFORM || F16) {
                    // pieces were split once per batch (split_tf32_kernel / split_f16_kernel): half 0 copies hi to columns [0, PIECE_COLS), half 1 lo behind it
                    // (wide tf32 rows: half 0 copies the whole hi piece, the lo piece goes to shared memory above)
                    const uint32_t kp = F16 ? p.nslab * 32u : p.nslab * SLAB_ELEMS;    // 32-bit words of a piece row (3xFP16: two elements per word)
                    const uint4* src = reinterpret_cast<const uint4*>(static_cast<const float*>(p.q_op) + (static_cast<uint64_t>(half) * p.nq + (has_query ? pr.x : 0)) * kp);
                    const uint32_t tq = tmem_base + ((quarter * 32u) << 16) + half * PIECE_COLS;
                    uint32_t c = 0;
                    for (; c + 64 <= kp; c += 64) {   // 16 loads in flight per round trip to L2
                        uint32_t w[64];
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            uint4 x = make_uint4(0u, 0u, 0u, 0u);
                            if (has_query) x = __ldg(src + c / 4 + j);
                            w[4 * j] = x.x; w[4 * j + 1] = x.y; w[4 * j + 2] = x.z; w[4 * j + 3] = x.w;
                        }
                        tmem_st64(tq + c, w);
                    }
                    if (c < kp) {
                        uint32_t w[32];
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            uint4 x = make_uint4(0u, 0u, 0u, 0u);
                            if (has_query) x = __ldg(src + c / 4 + j);
                            w[4 * j] = x.x; w[4 * j + 1] = x.y; w[4 * j + 2] = x.z; w[4 * j + 3] = x.w;
                        }
                        tmem_st32(tq + c, w);
                    }
                } else {
                    // bf16 terms q0, q1, q2 (split once per batch) at columns [0,64), [64,128), [128,192), two elements per column:
                    // half 0 copies q0 and the even 32-column chunks of q2, half 1 copies q1 and the odd chunks of q2
                    const uint32_t row_words = kp / 2;
                    auto copy_chunk = [&](uint32_t pc, uint32_t c) {
                        const uint4* src = reinterpret_cast<const uint4*>(static_cast<const uint32_t*>(p.q_op) +
                                                                          (static_cast<uint64_t>(pc) * p.nq + (has_query ? pr.x : 0)) * row_words + c);
                        uint32_t w[32];
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            uint4 x = make_uint4(0u, 0u, 0u, 0u);
                            if (has_query) x = __ldg(src + j);
                            w[4 * j] = x.x; w[4 * j + 1] = x.y; w[4 * j + 2] = x.z; w[4 * j + 3] = x.w;
                        }
                        tmem_st32(tmem_base + ((quarter * 32u) << 16) + pc * PIECE_COLS + c, w);
                    };
                    for (uint32_t c = 0; c < row_words; c += 32) copy_chunk(half, c);
                    if (p.bf16_terms > 2)
                        for (uint32_t c = half * 32; c < row_words; c += 64) copy_chunk(2, c);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(bar_q);
                c_gather += tc_clock() - c_t1;
            }
            top.init();
            // 3xFP16: undo the operands' power-of-two scales -- one per index for L2 lists and one per query, both in cq; cosine lists
            // keep per-row scales folded into aux
            constexpr bool RX = false;     // (L2 list operands carry one uniform scale: nothing per column to undo)
            float cq = 1.0f;
            if (F16) cq = (has_query ? __ldg(p.q_inv_scale + pr.x) : 1.0f) * ((MET == MET_L2) ? -2.0f * p.db_inv_scale : 1.0f);
            uint32_t* gtau_ptr = p.gtau + (has_query ? pr.x : 0);
            uint32_t g_next = has_query ? *reinterpret_cast<volatile uint32_t*>(gtau_ptr) : 0u;   // 0 = ordered(-NaN): prunes everything
            const float* aux_half = p.aux + r_begin + half * 64 + lane;
            auto load_aux = [&](uint32_t t, float& lo_v, float& hi_v) {
                const uint64_t c0 = r_begin + static_cast<uint64_t>(t) * BN + half * 64 + lane;
                lo_v = (c0 < r_end) ? __ldg(aux_half + static_cast<size_t>(t) * BN) : __int_as_float(0x7FC00000);        // NaN masks rows of other lists
                hi_v = (c0 + 32 < r_end) ? __ldg(aux_half + static_cast<size_t>(t) * BN + 32) : __int_as_float(0x7FC00000);
            };
            float* s_aux = s_aux_all + (warp - EPI_WARP0) * 64;   // warp-private: read back as broadcast loads by the value loop
            float* s_rx = s_aux_all + (8 + (warp - EPI_WARP0)) * 64;
            const float* rx_half = RX ? p.aux2 + r_begin + half * 64 + lane : nullptr;
            float aux_lo_next = 0.f, aux_hi_next = 0.f, rx_lo_next = 1.f, rx_hi_next = 1.f;
            if (n_tiles > 0) load_aux(0, aux_lo_next, aux_hi_next);
            if (RX && n_tiles > 0) { rx_lo_next = __ldg(rx_half); rx_hi_next = __ldg(rx_half + 32); }
            for (uint32_t t = 0; t < n_tiles; t++, next_tile()) {
                const uint32_t acc = r_acc, aph = r_aph;
                const uint32_t row0 = static_cast<uint32_t>(r_begin) + t * BN;
                const uint32_t g_bits = g_next;
                const float aux_lo = aux_lo_next, aux_hi = aux_hi_next;
                const float rx_lo = rx_lo_next, rx_hi = rx_hi_next;
                if (t + 1 < n_tiles) {
                    load_aux(t + 1, aux_lo_next, aux_hi_next);
                    if (RX) { rx_lo_next = __ldg(rx_half + static_cast<size_t>(t + 1) * BN); rx_hi_next = __ldg(rx_half + static_cast<size_t>(t + 1) * BN + 32); }
                    if (has_query) g_next = *reinterpret_cast<volatile uint32_t*>(gtau_ptr);
                }
                if (warp_has_query) {
                    __syncwarp();                  // every lane is done with the previous tile's constants
                    s_aux[lane] = aux_lo;
                    s_aux[lane + 32] = aux_hi;
                    if (RX) { s_rx[lane] = rx_lo; s_rx[lane + 32] = rx_hi; }
                    __syncwarp();
                }
                mbar_wait_timed(bar_tfull + acc, aph, c_tfull);
                tc_fence_after();
                const float g_tau = has_query ? ordered_to_f32(g_bits) : -INFINITY;   // lanes without a query never select
                float tau = fminf(top.tau(), g_tau);
                const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + ACC_COL0 + acc * BN;
                if (warp_has_query) {
                    const int c = static_cast<int>(half);
                    uint32_t r[64];
                    tmem_ld64_sync(taddr + c * 64, r);
                    tc_fence_before();             // values are in registers: free the accumulator stage before the select work
                    mbar_arrive(bar_tempty + acc);
                    float v[64];
                    float gm[8];
#pragma unroll
                    for (int g = 0; g < 8; g++) {
                        float mg = INFINITY;
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const int col = g * 8 + j;
                            const float cst = s_aux[col];
                            const float sdot = (KIND == KIND_I8) ? __int2float_rn(static_cast<int32_t>(r[col])) : __uint_as_float(r[col]);
                            if (F16) v[col] = (MET == MET_L2) ? fmaf(sdot, cq, cst) : (sdot * cst) * cq;
                            else v[col] = (MET == MET_L2) ? fmaf(sdot, -2.0f, cst) : sdot * cst;
                            mg = fminf(mg, v[col]);
                        }
                        gm[g] = mg;
                    }
                    const float m = fminf(fminf(fminf(gm[0], gm[1]), fminf(gm[2], gm[3])), fminf(fminf(gm[4], gm[5]), fminf(gm[6], gm[7])));
                    if (m < tau) select_from_tile<KP>(top, tau, v, gm, m, row0 + c * 64, scratch);
                } else {
                    tc_fence_before();
                    mbar_arrive(bar_tempty + acc);
                }
                if (has_query && (top.tau() < g_tau || (g_tau != g_tau && top.tau() < INFINITY))) atomicMin(gtau_ptr, f32_to_ordered(top.tau()));
            }
            if (has_query) {
                uint64_t* out = p.part_keys + ((static_cast<uint64_t>(pr.x) * p.probe_pitch + pr.y) * 2 + half) * KP;
#pragma unroll
                for (int j = 0; j < KP; j++) out[j] = (top.i[j] == IDX_INVALID) ? KEY_SENTINEL : make_key(top.v[j], top.i[j]);
            }
        }
        task_no++;
        c_tasks++;
        c_tiles += n_tiles;
    }
    if (dbg_on) {
        if (threadIdx.x == 0) { p.dbg[0] = tc_clock() - c_start; p.dbg[1] = c_sched; p.dbg[7] = (static_cast<unsigned long long>(c_tasks) << 32) | static_cast<unsigned long long>(c_tiles); }
        if (threadIdx.x == 32) { p.dbg[4] = c_wq; p.dbg[5] = c_wdata; p.dbg[6] = c_wtempty; }
        if (threadIdx.x == EPI_WARP0 * 32) { p.dbg[2] = c_gather; p.dbg[3] = c_tfull; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}


}  // namespace tc

// =============================================================================================== host side
// (EncodeTiled / tensor-map helpers live in flat_tc.cu)
int tc_make_tmap(CUtensorMap* tm, void* base, uint64_t rows, uint32_t kp_elems, int elem_bytes);
uint32_t tc_blocks_for(uint64_t work);
int tc_compute_xnorm_max(annb_index* ix, const float* d_aux, uint64_t n);

struct IvfTcState {
    int kind = -1;
    uint32_t kp_elems = 0, nslab = 0, n_pad = 0;
    void* d_x = nullptr;
    float* d_aux = nullptr;
    float* d_aux2 = nullptr;   // KIND_F16X3: inverse power-of-two operand scale of every stored row
    float db_inv_scale = 1.0f; // KIND_F16X3, L2: 1 / (uniform operand scale)
    CUtensorMap tm_x;
    DevBuf part, gtau, dbgc, q_op, q_scale;
    uint64_t bytes = 0;
};

int tc_ivf_prepare(annb_index* ix) {
    if (!ix->is_ivf || ix->n == 0) return ANNB_OK;
    int kind = ix->dtype == ANNB_F32 ? tc::KIND_TF32X3 : (ix->dtype == ANNB_BF16 ? tc::KIND_BF16 : tc::KIND_I8);
    // f32 lists with rows of up to 256 elements: 3xFP16 (see split_f16_kernel; the kernel converts the raw f32 slabs in place)
    if (kind == tc::KIND_TF32X3 && ix->opt_tc_f32_fp16 != 0 && round_up(ix->dim, 64u) <= 256u) kind = tc::KIND_F16X3;
    const uint32_t elem = kind == tc::KIND_TF32X3 ? 4 : (kind == tc::KIND_I8 ? 1 : 2);   // element size of the operand TMA loads
    const uint32_t slab_elems = tc::SLAB_BYTES / elem;                                  // K elements per ring stage
    const uint32_t kp = round_up(ix->dim, slab_elems);
    // the query pieces live in TMEM (128 columns per f32 / int8 piece, 64 or 128 per bf16 term); larger dims stay on the CUDA-core scan
    // (f32 rows of up to 1024 B keep their lo piece in shared memory, IvfTcParams::lo_smem)
    if (kp * elem > (kind == tc::KIND_TF32X3 ? 1024u : 512u)) return ANNB_OK;
    IvfTcState* st = new IvfTcState();
    ix->tc_ivf = st;
    st->kind = kind;
    st->kp_elems = kp;
    st->nslab = kp / slab_elems;
    st->n_pad = static_cast<uint32_t>(round_up<uint64_t>(ix->n, tc::BN)) + tc::BN;   // tiles may start anywhere: one extra tile of padding
    cudaStream_t s = ix->stream;
    const uint64_t aux_rows = static_cast<uint64_t>(st->n_pad) + tc::BN;
    {
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&st->d_aux), aux_rows * sizeof(float));
        if (e != cudaSuccess) { (void)cudaGetLastError(); set_last_error(std::string("cudaMalloc ivf tc aux: ") + cudaGetErrorString(e)); return ANNB_ERR_OUT_OF_MEMORY; }
        st->bytes += aux_rows * sizeof(float);
    }
    tc::aux_kernel<<<static_cast<uint32_t>((aux_rows + 127) / 128), 128, 0, s>>>(ix->d_rows, ix->row_bytes, kind == tc::KIND_F16X3 ? 0 : kind, ix->dim, ix->d_norms, ix->d_norms_i,
                                                                                ix->metric == ANNB_COSINE, ix->n, aux_rows, st->d_aux);
    ANNB_CUDA_CHECK(cudaGetLastError());
    if (kind == tc::KIND_F16X3) {
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&st->d_aux2), aux_rows * sizeof(float));
        if (e != cudaSuccess) { (void)cudaGetLastError(); set_last_error(std::string("cudaMalloc ivf tc row scales: ") + cudaGetErrorString(e)); return ANNB_ERR_OUT_OF_MEMORY; }
        st->bytes += aux_rows * sizeof(float);
        tc::fill_f32_value_kernel<<<static_cast<uint32_t>((aux_rows + 127) / 128), 128, 0, s>>>(st->d_aux2, aux_rows, 1.0f);
        const uint64_t xbytes = 2ull * st->n_pad * kp * 2;
        e = cudaMalloc(&st->d_x, xbytes);
        if (e != cudaSuccess) { (void)cudaGetLastError(); set_last_error(std::string("cudaMalloc ivf tc operand: ") + cudaGetErrorString(e)); return ANNB_ERR_OUT_OF_MEMORY; }
        st->bytes += xbytes;
        float sc = 0.0f;                   // L2: one uniform scale for the lists (0 = per-row scales, cosine)
        if (ix->metric != ANNB_COSINE) {
            ANNB_TRY(tc_uniform_f16_scale(ix, reinterpret_cast<const float*>(ix->d_rows), ix->n * static_cast<uint64_t>(ix->row_bytes / 4), &sc));
            st->db_inv_scale = 1.0f / sc;
        }
        tc::split_f16_kernel<<<tc_blocks_for(static_cast<uint64_t>(st->n_pad) * 32), 256, 0, s>>>(reinterpret_cast<const float*>(ix->d_rows), ix->row_bytes / 4, ix->dim, ix->n, st->n_pad, kp,
                                                                                            static_cast<__half*>(st->d_x), st->d_aux2, 1.0f, sc);
        if (ix->metric == ANNB_COSINE) tc::mul_rows_kernel<<<static_cast<uint32_t>((ix->n + 127) / 128), 128, 0, s>>>(st->d_aux, st->d_aux2, ix->n);
        ANNB_CUDA_CHECK(cudaGetLastError());
    }
    // Database operand: the index's own rows whenever their pitch is a whole number of 128-byte K slabs (dim % 32 == 0 for
    // f32, % 64 for bf16, % 128 for SQ8) -- TMA zero-fills the rows past the end; otherwise a zero-padded copy.  f32 rows
    // are split into tf32 hi / lo inside the kernel.
    void* xbase = ix->d_rows;
    if (kind == tc::KIND_F16X3) {
        xbase = st->d_x;
    } else if (ix->row_bytes != kp * elem) {
        const uint64_t bytes = static_cast<uint64_t>(ix->n) * kp * elem;
        cudaError_t e = cudaMalloc(&st->d_x, bytes);
        if (e != cudaSuccess) { (void)cudaGetLastError(); set_last_error(std::string("cudaMalloc ivf tc operand: ") + cudaGetErrorString(e)); return ANNB_ERR_OUT_OF_MEMORY; }
        st->bytes += bytes;
        tc::repitch_rows_kernel<<<tc_blocks_for(bytes), 256, 0, s>>>(ix->d_rows, ix->row_bytes, static_cast<uint8_t*>(st->d_x), kp * elem, ix->n);
        ANNB_CUDA_CHECK(cudaGetLastError());
        xbase = st->d_x;
    }
    ANNB_TRY(tc_make_tmap(&st->tm_x, xbase, kind == tc::KIND_F16X3 ? 2ull * st->n_pad : ix->n, kp, elem));
    ANNB_TRY(tc_compute_xnorm_max(ix, st->d_aux, ix->n));
    ix->device_bytes += st->bytes;
    return ANNB_OK;
}

void tc_ivf_destroy(annb_index* ix) {
    if (!ix->tc_ivf) return;
    ix->device_bytes -= std::min<uint64_t>(ix->device_bytes, ix->tc_ivf->bytes);
    cudaFree(ix->tc_ivf->d_x);
    cudaFree(ix->tc_ivf->d_aux);
    cudaFree(ix->tc_ivf->d_aux2);
    ix->tc_ivf->q_scale.release();
    ix->tc_ivf->part.release();
    ix->tc_ivf->gtau.release();
    ix->tc_ivf->dbgc.release();
    ix->tc_ivf->q_op.release();
    delete ix->tc_ivf;
    ix->tc_ivf = nullptr;
}

// Test hooks: role wait-cycle counters of CTA 0 of the scan kernel.
int tc_ivf_debug_enable(annb_index* ix, bool on) {
    if (!ix->tc_ivf) return ANNB_ERR_UNSUPPORTED;
    if (!on) { ix->tc_ivf->dbgc.release(); return ANNB_OK; }
    ANNB_TRY(ix->tc_ivf->dbgc.ensure(64));
    ANNB_CUDA_CHECK(cudaMemset(ix->tc_ivf->dbgc.p, 0, 64));
    return ANNB_OK;
}
int tc_ivf_debug_cycles(annb_index* ix, unsigned long long* host_out8) {
    if (!ix->tc_ivf || !ix->tc_ivf->dbgc.p) return ANNB_ERR_UNSUPPORTED;
    ANNB_CUDA_CHECK(cudaMemcpy(host_out8, ix->tc_ivf->dbgc.p, 64, cudaMemcpyDeviceToHost));
    return ANNB_OK;
}
// Test hook: the first 64 KiB of the per-(query, rank, half) candidate lists of the last scan (packed keys).
int tc_ivf_debug_fetch(annb_index* ix, float* host_out) {
    if (!ix->tc_ivf || !ix->tc_ivf->part.p) return ANNB_ERR_UNSUPPORTED;
    ANNB_CUDA_CHECK(cudaMemcpy(host_out, ix->tc_ivf->part.p, std::min<size_t>(65536, ix->tc_ivf->part.cap), cudaMemcpyDeviceToHost));
    return ANNB_OK;
}

int tc_ivf_kind(const annb_index* ix) { return ix->tc_ivf ? ix->tc_ivf->kind : -1; }

bool tc_ivf_supported(const annb_index* ix, int qt, uint32_t k_eff) {
    if (ix->tc_ivf == nullptr || k_eff > 24) return false;
    return ix->dtype == ANNB_SQ8 ? qt == QT_I8 : qt == QT_F32;
}

uint32_t tc_ivf_kprime(const annb_index* ix, uint32_t k_eff) {
    if (ix->opt_tc_candidates == 32) return 32;
    if (ix->tc_ivf && ix->tc_ivf->kind == tc::KIND_TF32X3 && ix->tc_ivf->kp_elems > 128) return 32;   // wide rows: wider certificate margin (tc_cert_eps)
    return k_eff <= 10 ? 16 : 32;
}

template <int KIND, int KP, int MET>
static int launch_ivf_tc(const CUtensorMap& tmx, const tc::IvfTcParams& p, uint32_t grid, size_t smem, cudaStream_t s) {
    auto kern = tc::ivf_tc_kernel<KIND, KP, MET>;
    ANNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, tc::ivf_tc_threads<KIND>(), smem, s>>>(tmx, p);
    ANNB_CUDA_CHECK(cudaGetLastError());
    return ANNB_OK;
}

template <int RT, int MET>
static int launch_ivf_rerank(const tc::RerankParams& r, cudaStream_t s) {
    auto kern = tc::rerank_kernel<RT, (RT == 2) ? QT_I8 : QT_F32, MET>;
    const size_t smem = static_cast<size_t>(r.nsort) * 8;
    if (smem > 200 * 1024) { set_last_error("ivf rerank: nprobe * k' too large"); return ANNB_ERR_UNSUPPORTED; }
    ANNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<static_cast<uint32_t>(r.nq), 128, smem, s>>>(r);
    ANNB_CUDA_CHECK(cudaGetLastError());
    return ANNB_OK;
}

// Scan + exact re-rank.  The pair buckets (grouped by list, 128 pairs per task) were built by the caller.
int tc_ivf_scan(annb_index* ix, const uint8_t* d_q, uint32_t q_bytes, uint64_t nq, uint32_t k_eff, uint32_t k_out, uint32_t probe_pitch,
                const uint32_t* d_pair_off, const uint32_t* d_task_off, const void* d_pairs, uint32_t* d_task_counter, uint64_t max_tasks,
                const uint32_t* d_n_probes, const uint64_t* row_map, uint64_t* d_ids, float* d_dist, uint32_t* d_cnt, cudaStream_t s,
                const void* d_tasks) {
    IvfTcState* st = ix->tc_ivf;
    const uint32_t kprime = tc_ivf_kprime(ix, k_eff);
    const uint32_t nb = (st->kind == tc::KIND_TF32X3 || st->kind == tc::KIND_F16X3) ? 2 : 1;
    const size_t fixed = 512 /*barriers, task slots (<= 296 B)*/ + 2 * 8 * 64 * 4 /*per-warp row constants, two banks*/;
    const size_t budget = 227 * 1024;
    const bool lo_s = st->kind == tc::KIND_TF32X3 && st->kp_elems > 128;
    const size_t q_smem = lo_s ? static_cast<size_t>(st->nslab) * tc::SLAB_TILE : 0;
    uint32_t stages = static_cast<uint32_t>(std::min<size_t>(8, (budget - fixed - q_smem) / (nb * tc::SLAB_TILE)));
    const size_t smem = q_smem + static_cast<size_t>(stages) * nb * tc::SLAB_TILE + fixed;
    const uint64_t slots = nq * static_cast<uint64_t>(probe_pitch) * 2;
    ANNB_TRY(st->part.ensure(slots * kprime * 8));
    ANNB_CUDA_CHECK(cudaMemsetAsync(st->part.p, 0xFF, slots * kprime * 8, s));
    ANNB_TRY(st->gtau.ensure(nq * 4 + 16));
    ANNB_CUDA_CHECK(cudaMemsetAsync(st->gtau.p, 0xFF, nq * 4 + 16, s));
    const bool l2 = ix->metric == ANNB_L2;
    // operand pieces of the batch's queries, split once (every query is gathered once per probed list)
    const uint32_t kp_q = st->kp_elems;
    const uint32_t bf16_terms = kp_q <= 128 ? tc_bf16_terms(ix) : 2u;   // three terms need rows of at most 128 elements (TMEM columns)
    if (st->kind == tc::KIND_F16X3) {
        ANNB_TRY(st->q_op.ensure(2ull * nq * kp_q * 2));
        ANNB_TRY(st->q_scale.ensure(nq * 4 + 16));
        tc::split_f16_kernel<<<tc_blocks_for(nq * 32), 256, 0, s>>>(reinterpret_cast<const float*>(d_q), q_bytes / 4, ix->dim, nq, nq, kp_q, st->q_op.as<__half>(), st->q_scale.as<float>());
        ANNB_CUDA_CHECK(cudaGetLastError());
        ix->stat_launches++;
    } else if (st->kind == tc::KIND_TF32X3) {
        ANNB_TRY(st->q_op.ensure(2ull * nq * kp_q * 4));
        tc::split_tf32_kernel<<<tc_blocks_for(nq * kp_q), 256, 0, s>>>(reinterpret_cast<const float*>(d_q), q_bytes / 4, ix->dim, nq, nq, kp_q, st->q_op.as<float>());
        ANNB_CUDA_CHECK(cudaGetLastError());
        ix->stat_launches++;
    } else if (st->kind == tc::KIND_BF16) {
        if (bf16_terms * (kp_q > 128 ? 128u : 64u) > 256u) { set_last_error("ivf tensor path: three bf16 query terms need rows of at most 128 elements"); return ANNB_ERR_UNSUPPORTED; }
        ANNB_TRY(st->q_op.ensure(3ull * nq * kp_q * 2));
        tc::split_bf16x3_kernel<<<tc_blocks_for(nq * kp_q), 256, 0, s>>>(reinterpret_cast<const float*>(d_q), q_bytes / 4, ix->dim, nq, nq, kp_q, st->q_op.as<__nv_bfloat16>());
        ANNB_CUDA_CHECK(cudaGetLastError());
        ix->stat_launches++;
    }
    tc::IvfTcParams p{};
    p.q_op = st->q_op.p; p.nq = nq; p.bf16_terms = bf16_terms; p.lo_smem = lo_s ? 1u : 0u; p.aux2 = st->d_aux2; p.q_inv_scale = st->q_scale.as<float>(); p.db_inv_scale = st->db_inv_scale;
    p.queries = d_q; p.q_bytes = q_bytes; p.dim = ix->dim; p.nslab = st->nslab; p.n_stages = stages; p.n_pad = st->n_pad; p.aux = st->d_aux;
    p.offsets = ix->d_offsets; p.shard_row0 = ix->shard_row0; p.nlist = ix->nlist; p.pair_off = d_pair_off; p.task_off = d_task_off;
    p.pairs = static_cast<const uint2*>(d_pairs); p.tasks = static_cast<const uint4*>(d_tasks); p.task_counter = d_task_counter; p.probe_pitch = probe_pitch; p.prefetch_task = ix->opt_ivf_task_prefetch ? 1u : 0u;
    p.part_keys = st->part.as<uint64_t>(); p.gtau = st->gtau.as<uint32_t>(); p.dbg = st->dbgc.as<unsigned long long>();
    const uint32_t grid = static_cast<uint32_t>(std::min<uint64_t>(std::max<uint64_t>(max_tasks, 1), 148));
    {
        cudaEvent_t ea = nullptr, eb = nullptr;
        if (ix->opt_time_kernels && cudaEventCreate(&ea) == cudaSuccess && cudaEventCreate(&eb) == cudaSuccess) cudaEventRecord(ea, s);
        int rc;
#define ANNB_IVF_TC(KIND_, KP_) (l2 ? launch_ivf_tc<KIND_, KP_, MET_L2>(st->tm_x, p, grid, smem, s) : launch_ivf_tc<KIND_, KP_, MET_COS>(st->tm_x, p, grid, smem, s))
        if (st->kind == tc::KIND_I8) rc = kprime == 16 ? ANNB_IVF_TC(tc::KIND_I8, 16) : ANNB_IVF_TC(tc::KIND_I8, 32);
        else if (st->kind == tc::KIND_F16X3) rc = kprime == 16 ? ANNB_IVF_TC(tc::KIND_F16X3, 16) : ANNB_IVF_TC(tc::KIND_F16X3, 32);
        else if (st->kind == tc::KIND_TF32X3) rc = kprime == 16 ? ANNB_IVF_TC(tc::KIND_TF32X3, 16) : ANNB_IVF_TC(tc::KIND_TF32X3, 32);
        else rc = kprime == 16 ? ANNB_IVF_TC(tc::KIND_BF16, 16) : ANNB_IVF_TC(tc::KIND_BF16, 32);
#undef ANNB_IVF_TC
        if (ea && eb) { cudaEventRecord(eb, s); ix->timed.emplace_back(ea, eb); }
        ANNB_TRY(rc);
        ix->stat_launches++;
    }
    tc::RerankParams r{};
    r.part_keys = st->part.as<uint64_t>(); r.parts = probe_pitch * 2; r.kp = kprime; r.k_eff = k_eff; r.k_out = k_out; r.gtau = st->gtau.as<uint32_t>();
    r.nsort = next_pow2(std::max(probe_pitch * 2 * kprime, 64u));
    r.nq = nq; r.rows = ix->d_rows; r.row_bytes = ix->row_bytes; r.row_norms = ix->d_norms; r.row_norms_i = ix->d_norms_i; r.queries = d_q; r.q_bytes = q_bytes; r.dim = ix->dim;
    r.bf16_self = 0; r.id_base = 0; r.parts_used = d_n_probes; r.part_mult = 2; r.id_map = ix->d_original_ids; r.row_map = row_map;
    r.out_ids = d_ids; r.out_dist = d_dist; r.out_counts = d_cnt;
    ANNB_TRY(ix->s_uncert.ensure((nq + 1) * 4));
    ANNB_CUDA_CHECK(cudaMemsetAsync(ix->s_uncert.p, 0, 4, s));
    r.cert_eps = tc_cert_eps(ix, st->kind, st->kp_elems, bf16_terms, st->kind == tc::KIND_TF32X3);
    { uint32_t b; std::memcpy(&b, &r.cert_eps, 4); ix->stat_cert_eps_bits = b; }
    r.xnorm_max = ix->tc_xnorm_max; r.uncert_count = ix->s_uncert.as<uint32_t>(); r.uncert_list = ix->s_uncert.as<uint32_t>() + 1; r.out_bound = ix->shard_bound;
    int rc;
    if (ix->dtype == ANNB_SQ8) rc = l2 ? launch_ivf_rerank<2, MET_L2>(r, s) : launch_ivf_rerank<2, MET_COS>(r, s);
    else if (ix->dtype == ANNB_F32) rc = l2 ? launch_ivf_rerank<0, MET_L2>(r, s) : launch_ivf_rerank<0, MET_COS>(r, s);
    else rc = l2 ? launch_ivf_rerank<1, MET_L2>(r, s) : launch_ivf_rerank<1, MET_COS>(r, s);
    ANNB_TRY(rc);
    ix->stat_launches++;
    return ANNB_OK;
}

}  // namespace annb
