// simt_kernels.cuh -- exact (reference-order) CUDA-core kernels.
//
//   tile_kernel     : a CTA holds 32 queries and streams 32-row tiles of a row store
//                     through shared memory (cp.async double buffer); every lane owns one
//                     row of the tile and computes its distance to 4 queries of its warp.
//                     Epilogues: top-k select (flat search), dense matrix (IVF centroid
//                     ranking), argmax (IVF coarse assignment at build time).
//   ivf_scan_kernel : one warp per (query, part) walks that query's probed inverted lists
//                     -- contiguous slabs in list order -- with 128-bit cp.async loads
//                     staged through shared memory, dequantisation fused into the distance.
//   probe_kernel    : per query, full (distance, cell) sort of the centroid ranking and the
//                     probe-expansion walk of select_probed_clusters.
//   finalize_kernel : merges per-split / per-part partial results into the final rows.
//
// All arithmetic goes through refdist.cuh, so results are bit-identical to the CPU
// reference's AVX2 path (see oracle/oracle.c).
#pragma once
#include "common.cuh"
#include "refdist.cuh"
#include "select.cuh"

namespace annb {

constexpr int TILE_ROWS = 32;      // rows per staged tile = one row per lane
constexpr int CTA_QUERIES = 32;    // queries per CTA in tile_kernel
constexpr int WARP_QUERIES = 4;    // queries per warp in tile_kernel
constexpr int TILE_THREADS = 256;

enum { EPI_SELECT = 0, EPI_DENSE = 1, EPI_ARGMAX = 2 };

// Shared-memory row pitch: an odd number of 16-byte chunks, so that the 8 lanes of a
// quarter warp reading the same chunk of 8 consecutive rows hit 8 distinct bank groups.
__host__ __device__ __forceinline__ uint32_t smem_row_pitch(uint32_t row_bytes) { return ((row_bytes >> 4) | 1u) << 4; }

struct TileParams {
    const uint8_t* rows;        // [n_rows][row_bytes]
    uint64_t n_rows;
    uint32_t row_bytes;
    const float* row_norms;     // cosine (f32 / bf16 rows)
    const int32_t* row_norms_i; // cosine (sq8 rows)
    const float* row_aux;       // EPI_ARGMAX: |c|^2 (L2) or 1/|c| (cosine) per centroid
    const uint8_t* queries;     // [nq][q_bytes]
    uint32_t q_bytes;
    uint64_t nq;
    uint32_t dim;
    int bf16_self;              // bf16 x bf16: query norm is rounded to bf16 (exhaustive_bf16.rs:259-270)
    // EPI_SELECT
    uint32_t k, nsort, n_splits;
    uint64_t rows_per_split;
    uint64_t* part_keys;        // [nq][n_splits][k]
    // EPI_DENSE
    float* dense_out;           // [nq][n_rows]
    // EPI_ARGMAX
    uint32_t* assign_out;       // [nq]
    int assign_cosine;
};

// Cooperative copy of `nrows` consecutive rows into a staged tile.
__device__ __forceinline__ void stage_rows(uint8_t* smem_tile, uint32_t pitch, const uint8_t* gsrc, uint32_t row_bytes,
                                           uint32_t nrows, uint32_t tid, uint32_t nthreads) {
    const uint32_t cpr = row_bytes >> 4;
    const uint32_t total = nrows * cpr;
    for (uint32_t g = tid; g < total; g += nthreads) {
        uint32_t r = g / cpr, c = g - r * cpr;
        cp_async16(smem_tile + r * pitch + (c << 4), gsrc + (static_cast<uint64_t>(g) << 4));
    }
}

// RT: 0 = f32 rows, 1 = bf16 rows, 2 = int8 rows.  QT: query element type (QT_*).
template <int RT, int QT, int MET, int EPI>
__global__ void __launch_bounds__(TILE_THREADS) tile_kernel(TileParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t xpitch = smem_row_pitch(p.row_bytes);
    uint8_t* s_q = smem;                                              // [32][q_bytes]
    float* s_qnorm = reinterpret_cast<float*>(s_q + CTA_QUERIES * p.q_bytes);  // [32]
    int32_t* s_qnsq = reinterpret_cast<int32_t*>(s_qnorm + CTA_QUERIES);       // [32]
    uint8_t* s_x = reinterpret_cast<uint8_t*>(s_qnsq + CTA_QUERIES);           // [2][32][xpitch]
    uint64_t* s_sel = reinterpret_cast<uint64_t*>(s_x + 2 * TILE_ROWS * xpitch);  // [32][nsort]

    const uint64_t q0 = static_cast<uint64_t>(blockIdx.x) * CTA_QUERIES;
    uint64_t r_begin = 0, r_end = p.n_rows;
    if (EPI == EPI_SELECT) {
        r_begin = static_cast<uint64_t>(blockIdx.y) * p.rows_per_split;
        r_end = min(p.n_rows, r_begin + p.rows_per_split);
        if (r_begin > r_end) r_begin = r_end;
    }
    // ---- queries -> shared (zero rows for q >= nq) ----
    {
        const uint32_t cpr = p.q_bytes >> 4;
        for (uint32_t g = tid; g < CTA_QUERIES * cpr; g += TILE_THREADS) {
            uint32_t r = g / cpr, c = g - r * cpr;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (q0 + r < p.nq) v = *reinterpret_cast<const uint4*>(p.queries + (q0 + r) * p.q_bytes + (c << 4));
            *reinterpret_cast<uint4*>(s_q + r * p.q_bytes + (c << 4)) = v;
        }
    }
    __syncthreads();
    if (tid < CTA_QUERIES) {
        const uint8_t* q = s_q + tid * p.q_bytes;
        float qn = 1.0f;
        int32_t qs = 0;
        if (QT == QT_I8) {
            for (uint32_t e = 0; e < p.dim; e++) {
                int32_t v = reinterpret_cast<const int8_t*>(q)[e];
                qs += v * v;
            }
        } else if (MET == MET_COS) {
            qn = seq_norm<(QT == QT_F32) ? 4 : 2>(q, p.dim);
            if (p.bf16_self) qn = round_to_bf16(qn);
        }
        s_qnorm[tid] = qn;
        s_qnsq[tid] = qs;
    }
    WarpSelect sel[WARP_QUERIES];
    if (EPI == EPI_SELECT) {
#pragma unroll
        for (int i = 0; i < WARP_QUERIES; i++)
            sel[i].init(s_sel + static_cast<size_t>(warp * WARP_QUERIES + i) * p.nsort, p.nsort, p.k);
    }
    float best_s[WARP_QUERIES];
    uint32_t best_c[WARP_QUERIES];
#pragma unroll
    for (int i = 0; i < WARP_QUERIES; i++) { best_s[i] = -INFINITY; best_c[i] = 0; }
    __syncthreads();

    const uint8_t* qw = s_q + warp * WARP_QUERIES * p.q_bytes;
    const uint64_t n_tiles = (r_end - r_begin + TILE_ROWS - 1) / TILE_ROWS;
    if (n_tiles > 0) {
        uint32_t nr = static_cast<uint32_t>(min(static_cast<uint64_t>(TILE_ROWS), r_end - r_begin));
        stage_rows(s_x, xpitch, p.rows + r_begin * p.row_bytes, p.row_bytes, nr, tid, TILE_THREADS);
    }
    cp_async_commit();
    for (uint64_t t = 0; t < n_tiles; t++) {
        const uint64_t row0 = r_begin + t * TILE_ROWS;
        if (t + 1 < n_tiles) {
            const uint64_t nrow0 = row0 + TILE_ROWS;
            uint32_t nr = static_cast<uint32_t>(min(static_cast<uint64_t>(TILE_ROWS), r_end - nrow0));
            stage_rows(s_x + ((t + 1) & 1) * TILE_ROWS * xpitch, xpitch, p.rows + nrow0 * p.row_bytes, p.row_bytes, nr, tid,
                       TILE_THREADS);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const uint64_t row = row0 + lane;
        const bool valid = row < r_end;
        const uint8_t* xr = s_x + (t & 1) * TILE_ROWS * xpitch + lane * xpitch;
        float dist[WARP_QUERIES];
#pragma unroll
        for (int i = 0; i < WARP_QUERIES; i++) dist[i] = 0.0f;
        if (valid) {
            if (RT == 2) {
                int32_t dot[WARP_QUERIES], xx;
                accumulate_i8<WARP_QUERIES>(xr, qw, p.q_bytes, p.dim, dot, xx);
                int32_t xn = (MET == MET_COS) ? p.row_norms_i[row] : 0;
#pragma unroll
                for (int i = 0; i < WARP_QUERIES; i++)
                    dist[i] = finish_i8<MET>(dot[i], xx, s_qnsq[warp * WARP_QUERIES + i], xn);
            } else {
                float raw[WARP_QUERIES];
                accumulate_fp<(RT == 0) ? 4 : 2, (QT == QT_F32) ? 4 : 2, MET == MET_L2, WARP_QUERIES>(xr, qw, p.q_bytes, p.dim,
                                                                                                     raw);
                float xn = (MET == MET_COS) ? p.row_norms[row] : 1.0f;
#pragma unroll
                for (int i = 0; i < WARP_QUERIES; i++) dist[i] = finish_fp<MET>(raw[i], s_qnorm[warp * WARP_QUERIES + i], xn);
            }
        }
        if (EPI == EPI_SELECT) {
#pragma unroll
            for (int i = 0; i < WARP_QUERIES; i++) sel[i].offer(make_key(dist[i], static_cast<uint32_t>(row)), valid);
        } else if (EPI == EPI_DENSE) {
            if (valid) {
#pragma unroll
                for (int i = 0; i < WARP_QUERIES; i++) {
                    uint64_t q = q0 + warp * WARP_QUERIES + i;
                    if (q < p.nq) p.dense_out[q * p.n_rows + row] = dist[i];
                }
            }
        } else {  // EPI_ARGMAX: direct_assign, src/utils/k_means_utils.rs:2119-2195
            if (valid) {
                float aux = p.row_aux[row];
#pragma unroll
                for (int i = 0; i < WARP_QUERIES; i++) {
                    float score = p.assign_cosine ? __fmul_rn(dist[i], aux) : __fsub_rn(__fmul_rn(2.0f, dist[i]), aux);
                    if (score > best_s[i]) { best_s[i] = score; best_c[i] = static_cast<uint32_t>(row); }
                }
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
    if (EPI == EPI_SELECT) {
#pragma unroll
        for (int i = 0; i < WARP_QUERIES; i++) {
            sel[i].flush();
            uint64_t q = q0 + warp * WARP_QUERIES + i;
            if (q < p.nq) {
                uint64_t* out = p.part_keys + (q * p.n_splits + blockIdx.y) * p.k;
                for (uint32_t j = lane; j < p.k; j += 32) out[j] = sel[i].buf[j];
            }
        }
    } else if (EPI == EPI_ARGMAX) {
#pragma unroll
        for (int i = 0; i < WARP_QUERIES; i++) {
            float s = best_s[i];
            uint32_t c = best_c[i];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                float os = __shfl_xor_sync(0xFFFFFFFFu, s, off);
                uint32_t oc = __shfl_xor_sync(0xFFFFFFFFu, c, off);
                if (os > s || (os == s && oc < c)) { s = os; c = oc; }
            }
            uint64_t q = q0 + warp * WARP_QUERIES + i;
            if (lane == 0 && q < p.nq) p.assign_out[q] = c;
        }
    }
}

static inline size_t tile_kernel_smem(uint32_t row_bytes, uint32_t q_bytes, uint32_t nsort, bool select) {
    size_t s = static_cast<size_t>(CTA_QUERIES) * q_bytes + CTA_QUERIES * 8 + 2ull * TILE_ROWS * smem_row_pitch(row_bytes);
    if (select) s += static_cast<size_t>(CTA_QUERIES) * nsort * 8;
    return s;
}

// ---------------------------------------------------------------------------
// IVF list scan (src/cpu/ivf.rs:367-381).
// ---------------------------------------------------------------------------
constexpr int SCAN_WARPS = 4;

struct ScanParams {
    const uint8_t* rows;         // this shard's vectors, list order
    uint32_t row_bytes;
    const float* row_norms;
    const int32_t* row_norms_i;
    const uint8_t* queries;      // scan-side queries [nq][q_bytes]
    uint32_t q_bytes;
    uint64_t nq;
    uint32_t dim;
    int bf16_self;
    const uint32_t* probes;      // [nq][probe_pitch] cell ids in rank order
    uint32_t probe_pitch;
    const uint32_t* n_probes;    // [nq]
    const uint64_t* offsets;     // [nlist+1] global CSR offsets
    uint32_t list_begin, list_end;
    uint64_t shard_row0;         // offsets[list_begin]
    uint32_t parts, k, nsort;
    uint32_t subs;               // every probed list is cut into `subs` row segments, one warp each (tiny batches)
    uint64_t* part_keys;         // [nq][parts * subs][k]; key index = row position inside the shard
};

template <int RT, int QT, int MET>
__global__ void __launch_bounds__(SCAN_WARPS * 32) ivf_scan_kernel(ScanParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t xpitch = smem_row_pitch(p.row_bytes);
    const size_t per_warp = static_cast<size_t>(p.q_bytes) + 2ull * TILE_ROWS * xpitch + static_cast<size_t>(p.nsort) * 8;
    uint8_t* base = smem + warp * per_warp;
    uint64_t* s_sel = reinterpret_cast<uint64_t*>(base);                 // [nsort]   (8-byte aligned first)
    uint8_t* s_q = base + static_cast<size_t>(p.nsort) * 8;              // [q_bytes]
    uint8_t* s_x = s_q + p.q_bytes;                                       // [2][32][xpitch]

    const uint64_t task = static_cast<uint64_t>(blockIdx.x) * SCAN_WARPS + warp;
    const uint32_t slices = p.parts * p.subs;
    const uint64_t q = task / slices;
    const uint32_t slice = static_cast<uint32_t>(task % slices);
    const uint32_t part = slice / p.subs, sub = slice - part * p.subs;
    if (q >= p.nq) return;  // whole warp exits together (task is warp-uniform)

    for (uint32_t c = lane; c < (p.q_bytes >> 4); c += 32)
        *reinterpret_cast<uint4*>(s_q + (c << 4)) = *reinterpret_cast<const uint4*>(p.queries + q * p.q_bytes + (c << 4));
    __syncwarp();
    float qnorm = 1.0f;
    int32_t qnsq = 0;
    if (QT == QT_I8) {
        for (uint32_t e = lane; e < p.dim; e += 32) {
            int32_t v = reinterpret_cast<const int8_t*>(s_q)[e];
            qnsq += v * v;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) qnsq += __shfl_xor_sync(0xFFFFFFFFu, qnsq, off);
    } else if (MET == MET_COS) {
        if (lane == 0) {
            qnorm = seq_norm<(QT == QT_F32) ? 4 : 2>(s_q, p.dim);
            if (p.bf16_self) qnorm = round_to_bf16(qnorm);
        }
        qnorm = __shfl_sync(0xFFFFFFFFu, qnorm, 0);
    }
    WarpSelect sel;
    sel.init(s_sel, p.nsort, p.k);

    const uint32_t np = p.n_probes[q];
    const uint32_t* probes = p.probes + q * p.probe_pitch;

    // tile generator over this part's probe ranks (part, part + parts, ...)
    uint32_t rank = part;
    uint64_t pos = 0, end = 0;  // current list: shard-local row range [pos, end)
    auto next_tile = [&](uint64_t& t_row, uint32_t& t_n) -> bool {
        while (pos >= end) {
            if (rank >= np) return false;
            uint32_t c = probes[rank];
            rank += p.parts;
            if (c < p.list_begin || c >= p.list_end) continue;  // list lives on another shard
            pos = p.offsets[c] - p.shard_row0;
            end = p.offsets[c + 1] - p.shard_row0;
            if (p.subs > 1) {   // this warp's segment of the list (whole tiles)
                const uint64_t seg = ((end - pos + p.subs - 1) / p.subs + TILE_ROWS - 1) / TILE_ROWS * TILE_ROWS;
                pos = min(end, pos + sub * seg);
                end = min(end, pos + seg);
            }
        }
        t_row = pos;
        t_n = static_cast<uint32_t>(min(static_cast<uint64_t>(TILE_ROWS), end - pos));
        pos += t_n;
        return true;
    };

    uint64_t cur_row = 0, nxt_row = 0;
    uint32_t cur_n = 0, nxt_n = 0;
    bool have = next_tile(cur_row, cur_n);
    uint32_t b = 0;
    if (have) stage_rows(s_x, xpitch, p.rows + cur_row * p.row_bytes, p.row_bytes, cur_n, lane, 32);
    cp_async_commit();
    while (have) {
        bool have_next = next_tile(nxt_row, nxt_n);
        if (have_next)
            stage_rows(s_x + (b ^ 1u) * TILE_ROWS * xpitch, xpitch, p.rows + nxt_row * p.row_bytes, p.row_bytes, nxt_n, lane, 32);
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const bool valid = lane < cur_n;
        const uint64_t row = cur_row + lane;
        const uint8_t* xr = s_x + b * TILE_ROWS * xpitch + lane * xpitch;
        float dist = 0.0f;
        if (valid) {
            if (RT == 2) {
                int32_t dot[1], xx;
                accumulate_i8<1>(xr, s_q, p.q_bytes, p.dim, dot, xx);
                int32_t xn = (MET == MET_COS) ? p.row_norms_i[row] : 0;
                dist = finish_i8<MET>(dot[0], xx, qnsq, xn);
            } else {
                float raw[1];
                accumulate_fp<(RT == 0) ? 4 : 2, (QT == QT_F32) ? 4 : 2, MET == MET_L2, 1>(xr, s_q, p.q_bytes, p.dim, raw);
                float xn = (MET == MET_COS) ? p.row_norms[row] : 1.0f;
                dist = finish_fp<MET>(raw[0], qnorm, xn);
            }
        }
        sel.offer(make_key(dist, static_cast<uint32_t>(row)), valid);
        __syncwarp();
        have = have_next;
        cur_row = nxt_row;
        cur_n = nxt_n;
        b ^= 1u;
    }
    cp_async_wait<0>();
    sel.flush();
    uint64_t* out = p.part_keys + (q * slices + slice) * p.k;
    for (uint32_t j = lane; j < p.k; j += 32) out[j] = sel.buf[j];
}

static inline size_t scan_kernel_smem(uint32_t row_bytes, uint32_t q_bytes, uint32_t nsort) {
    return SCAN_WARPS * (static_cast<size_t>(q_bytes) + 2ull * TILE_ROWS * smem_row_pitch(row_bytes) + static_cast<size_t>(nsort) * 8);
}

// ---------------------------------------------------------------------------
// IVF list scan, list-major ("batched") variant.
//
// With a query batch every inverted list is probed by many queries (10k queries x 32 probes over 4096 lists: ~78
// queries per list).  The (query, probe-rank) pairs are bucketed by list on the device; one task = one list x a group
// of up to 32 of the queries that probe it.  A CTA stages the list's 32-row tiles once (cp.async double buffer) and
// every warp scores its 4 queries of the group against the staged rows -- the same inner loop as tile_kernel, same
// reference-order arithmetic -- so each byte fetched from HBM / L2 is used by up to 32 queries instead of one.
// Persistent CTAs pull tasks from an atomic counter (lists differ a lot in length).
// Output: one k-list per (query, probe rank), merged by finalize_kernel.
// ---------------------------------------------------------------------------
struct PairParams {
    const uint32_t* probes;    // [nq][probe_pitch]
    uint32_t probe_pitch;
    const uint32_t* n_probes;  // [nq]
    uint64_t nq;
    uint32_t nlist, list_begin, list_end;
    uint32_t* cnt;             // [nlist]   pairs per list (zeroed by the caller)
    uint32_t* cursor;          // [nlist]
    uint32_t* pair_off;        // [nlist+1]
    uint32_t* task_off;        // [nlist+1] prefix sum of ceil(cnt / group)
    uint32_t group;            // queries per task (32 for the CUDA-core kernel, 128 for the tensor-core kernel)
    uint2* pairs;              // [sum cnt] (query, rank) grouped by list
    uint32_t* task_counter;    // persistent-CTA work counter (zeroed by the caller)
    uint4* tasks;              // optional [2 * n_tasks]: ready-made task records {list, pair0, pairs in group, 0} {row begin, row end (u64 each)}
    const uint64_t* offsets;   // global CSR offsets (task records)
    uint64_t shard_row0;
    const uint32_t* order;     // optional [nlist]: the order in which lists receive their task slots (task records only)
    unsigned long long* stat_tiles;  // optional: receives the number of 128-row tiles the tasks will execute (sum over tasks of ceil(list rows / 128))
};

__global__ void ivf_count_pairs_kernel(PairParams p) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= p.nq * p.probe_pitch) return;
    const uint64_t q = i / p.probe_pitch;
    const uint32_t r = static_cast<uint32_t>(i - q * p.probe_pitch);
    if (r >= p.n_probes[q]) return;
    const uint32_t c = p.probes[i];
    if (c >= p.list_begin && c < p.list_end) atomicAdd(p.cnt + c, 1u);
}
// Exclusive scan of one value per thread over a block of 1024 threads (warp shuffles + one shared row of warp totals);
// *total receives the block sum.  Every thread of the block must call it.
__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t* s_warp /*[33]*/) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= static_cast<uint32_t>(off)) incl += t;
    }
    __syncthreads();                       // s_warp may still be read from a previous call
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t wi = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, off);
            if (lane >= static_cast<uint32_t>(off)) wi += t;
        }
        s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    return incl - v + s_warp[warp];
}
// Single block: exclusive scans of the per-list pair counts and task counts (nlist <= 16384).
__global__ void __launch_bounds__(1024) ivf_pair_offsets_kernel(PairParams p) {
    __shared__ uint32_t s_warp[33];
    const uint32_t per = (p.nlist + 1023) / 1024;
    const uint32_t lo = min(p.nlist, threadIdx.x * per), hi = min(p.nlist, lo + per);
    uint32_t a = 0, b = 0;
    for (uint32_t c = lo; c < hi; c++) { a += p.cnt[c]; b += (p.cnt[c] + p.group - 1) / p.group; }
    a = block_exclusive_scan_1024(a, s_warp);
    if (threadIdx.x == 0) p.pair_off[p.nlist] = s_warp[32];
    b = block_exclusive_scan_1024(b, s_warp);
    if (threadIdx.x == 0) p.task_off[p.nlist] = s_warp[32];
    if (p.stat_tiles != nullptr) {
        unsigned long long t = 0;
        for (uint32_t c = lo; c < hi; c++) t += static_cast<unsigned long long>((p.cnt[c] + p.group - 1) / p.group) * ((p.offsets[c + 1] - p.offsets[c] + 127) / 128);
        if (t) atomicAdd(p.stat_tiles, t);
    }
    const bool reorder = p.tasks != nullptr && p.order != nullptr;
    for (uint32_t c = lo; c < hi; c++) {
        p.pair_off[c] = a; p.cursor[c] = a; p.task_off[c] = b;
        const uint32_t nt = (p.cnt[c] + p.group - 1) / p.group;
        if (p.tasks != nullptr && !reorder) {   // one record per task, so the scan kernel fetches a task with a single round trip
            const uint64_t rb = p.offsets[c] - p.shard_row0, re = p.offsets[c + 1] - p.shard_row0;
            for (uint32_t g = 0; g < nt; g++) {
                p.tasks[2 * (b + g)] = make_uint4(c, a + g * p.group, min(p.group, p.cnt[c] - g * p.group), 0u);
                p.tasks[2 * (b + g) + 1] = make_uint4(static_cast<uint32_t>(rb), static_cast<uint32_t>(rb >> 32), static_cast<uint32_t>(re), static_cast<uint32_t>(re >> 32));
            }
        }
        a += p.cnt[c]; b += nt;
    }
    if (!reorder) return;
    // Task slots in the caller's list order (longest list first): a second scan of the task counts over the positions of
    // `order`; the pair offsets written above stay in list order.
    b = 0;
    for (uint32_t i = lo; i < hi; i++) b += (p.cnt[p.order[i]] + p.group - 1) / p.group;
    b = block_exclusive_scan_1024(b, s_warp);   // (its barriers also make pair_off[] of every list visible)
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t c = p.order[i];
        const uint32_t n_c = p.cnt[c], nt = (n_c + p.group - 1) / p.group, a0 = p.pair_off[c];
        const uint64_t rb = p.offsets[c] - p.shard_row0, re = p.offsets[c + 1] - p.shard_row0;
        for (uint32_t g = 0; g < nt; g++) {
            p.tasks[2 * (b + g)] = make_uint4(c, a0 + g * p.group, min(p.group, n_c - g * p.group), 0u);
            p.tasks[2 * (b + g) + 1] = make_uint4(static_cast<uint32_t>(rb), static_cast<uint32_t>(rb >> 32), static_cast<uint32_t>(re), static_cast<uint32_t>(re >> 32));
        }
        b += nt;
    }
}
__global__ void ivf_fill_pairs_kernel(PairParams p) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= p.nq * p.probe_pitch) return;
    const uint64_t q = i / p.probe_pitch;
    const uint32_t r = static_cast<uint32_t>(i - q * p.probe_pitch);
    if (r >= p.n_probes[q]) return;
    const uint32_t c = p.probes[i];
    if (c >= p.list_begin && c < p.list_end) p.pairs[atomicAdd(p.cursor + c, 1u)] = make_uint2(static_cast<uint32_t>(q), r);
}

struct ListScanParams {
    const uint8_t* rows;
    uint32_t row_bytes;
    const float* row_norms;
    const int32_t* row_norms_i;
    const uint8_t* queries;
    uint32_t q_bytes;
    uint32_t dim;
    int bf16_self;
    const uint64_t* offsets;   // global CSR offsets
    uint64_t shard_row0;
    uint32_t nlist;
    const uint32_t* pair_off;
    const uint32_t* task_off;
    const uint2* pairs;
    uint32_t* task_counter;
    uint32_t probe_pitch, k, nsort;
    uint64_t* part_keys;       // [nq][probe_pitch][k]
};

template <int RT, int QT, int MET>
__global__ void __launch_bounds__(TILE_THREADS) ivf_list_kernel(ListScanParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t xpitch = smem_row_pitch(p.row_bytes);
    uint8_t* s_q = smem;
    float* s_qnorm = reinterpret_cast<float*>(s_q + CTA_QUERIES * p.q_bytes);
    int32_t* s_qnsq = reinterpret_cast<int32_t*>(s_qnorm + CTA_QUERIES);
    uint2* s_pair = reinterpret_cast<uint2*>(s_qnsq + CTA_QUERIES);             // [32] (query, rank) of the group
    uint8_t* s_x = reinterpret_cast<uint8_t*>(s_pair + CTA_QUERIES);
    uint64_t* s_sel = reinterpret_cast<uint64_t*>(s_x + 2 * TILE_ROWS * xpitch);
    __shared__ uint32_t s_task;
    const uint32_t total_tasks = p.task_off[p.nlist];

    for (;;) {
        __syncthreads();  // previous task fully retired (shared buffers reusable)
        if (tid == 0) s_task = atomicAdd(p.task_counter, 1u);
        __syncthreads();
        const uint32_t task = s_task;
        if (task >= total_tasks) break;
        // list of this task: largest c with task_off[c] <= task
        uint32_t lo = 0, hi = p.nlist;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (p.task_off[mid] <= task) lo = mid; else hi = mid;
        }
        const uint32_t c = lo;
        const uint32_t g = task - p.task_off[c];
        const uint32_t pair0 = p.pair_off[c] + g * CTA_QUERIES;
        const uint32_t n_in_group = min(static_cast<uint32_t>(CTA_QUERIES), p.pair_off[c + 1] - pair0);
        const uint64_t r_begin = p.offsets[c] - p.shard_row0, r_end = p.offsets[c + 1] - p.shard_row0;

        if (tid < CTA_QUERIES) s_pair[tid] = (tid < n_in_group) ? p.pairs[pair0 + tid] : make_uint2(0xFFFFFFFFu, 0u);
        __syncthreads();
        {
            const uint32_t cpr = p.q_bytes >> 4;
            for (uint32_t i = tid; i < CTA_QUERIES * cpr; i += TILE_THREADS) {
                const uint32_t r = i / cpr, ch = i - r * cpr;
                uint4 v = make_uint4(0, 0, 0, 0);
                const uint32_t q = s_pair[r].x;
                if (q != 0xFFFFFFFFu) v = *reinterpret_cast<const uint4*>(p.queries + static_cast<uint64_t>(q) * p.q_bytes + (ch << 4));
                *reinterpret_cast<uint4*>(s_q + r * p.q_bytes + (ch << 4)) = v;
            }
        }
        __syncthreads();
        if (tid < CTA_QUERIES) {
            const uint8_t* q = s_q + tid * p.q_bytes;
            float qn = 1.0f;
            int32_t qs = 0;
            if (QT == QT_I8) {
                for (uint32_t e = 0; e < p.dim; e++) {
                    const int32_t v = reinterpret_cast<const int8_t*>(q)[e];
                    qs += v * v;
                }
            } else if (MET == MET_COS) {
                qn = seq_norm<(QT == QT_F32) ? 4 : 2>(q, p.dim);
                if (p.bf16_self) qn = round_to_bf16(qn);
            }
            s_qnorm[tid] = qn;
            s_qnsq[tid] = qs;
        }
        const bool warp_active = warp * WARP_QUERIES < n_in_group;
        WarpSelect sel[WARP_QUERIES];
#pragma unroll
        for (int i = 0; i < WARP_QUERIES; i++) sel[i].init(s_sel + static_cast<size_t>(warp * WARP_QUERIES + i) * p.nsort, p.nsort, p.k);
        __syncthreads();

        const uint8_t* qw = s_q + warp * WARP_QUERIES * p.q_bytes;
        const uint64_t n_tiles = (r_end - r_begin + TILE_ROWS - 1) / TILE_ROWS;
        if (n_tiles > 0) {
            const uint32_t nr = static_cast<uint32_t>(min(static_cast<uint64_t>(TILE_ROWS), r_end - r_begin));
            stage_rows(s_x, xpitch, p.rows + r_begin * p.row_bytes, p.row_bytes, nr, tid, TILE_THREADS);
        }
        cp_async_commit();
        for (uint64_t t = 0; t < n_tiles; t++) {
            const uint64_t row0 = r_begin + t * TILE_ROWS;
            if (t + 1 < n_tiles) {
                const uint64_t nrow0 = row0 + TILE_ROWS;
                const uint32_t nr = static_cast<uint32_t>(min(static_cast<uint64_t>(TILE_ROWS), r_end - nrow0));
                stage_rows(s_x + ((t + 1) & 1) * TILE_ROWS * xpitch, xpitch, p.rows + nrow0 * p.row_bytes, p.row_bytes, nr, tid, TILE_THREADS);
            }
            cp_async_commit();
            cp_async_wait<1>();
            __syncthreads();
            if (warp_active) {
                const uint64_t row = row0 + lane;
                const bool valid = row < r_end;
                const uint8_t* xr = s_x + (t & 1) * TILE_ROWS * xpitch + lane * xpitch;
                float dist[WARP_QUERIES];
#pragma unroll
                for (int i = 0; i < WARP_QUERIES; i++) dist[i] = 0.0f;
                if (valid) {
                    if (RT == 2) {
                        int32_t dot[WARP_QUERIES], xx;
                        accumulate_i8<WARP_QUERIES>(xr, qw, p.q_bytes, p.dim, dot, xx);
                        const int32_t xn = (MET == MET_COS) ? p.row_norms_i[row] : 0;
#pragma unroll
                        for (int i = 0; i < WARP_QUERIES; i++) dist[i] = finish_i8<MET>(dot[i], xx, s_qnsq[warp * WARP_QUERIES + i], xn);
                    } else {
                        float raw[WARP_QUERIES];
                        accumulate_fp<(RT == 0) ? 4 : 2, (QT == QT_F32) ? 4 : 2, MET == MET_L2, WARP_QUERIES>(xr, qw, p.q_bytes, p.dim, raw);
                        const float xn = (MET == MET_COS) ? p.row_norms[row] : 1.0f;
#pragma unroll
                        for (int i = 0; i < WARP_QUERIES; i++) dist[i] = finish_fp<MET>(raw[i], s_qnorm[warp * WARP_QUERIES + i], xn);
                    }
                }
#pragma unroll
                for (int i = 0; i < WARP_QUERIES; i++) sel[i].offer(make_key(dist[i], static_cast<uint32_t>(row)), valid);
            }
            __syncthreads();
        }
        cp_async_wait<0>();
        if (warp_active) {
#pragma unroll
            for (int i = 0; i < WARP_QUERIES; i++) {
                sel[i].flush();
                const uint2 pr = s_pair[warp * WARP_QUERIES + i];
                if (pr.x != 0xFFFFFFFFu) {
                    uint64_t* out = p.part_keys + (static_cast<uint64_t>(pr.x) * p.probe_pitch + pr.y) * p.k;
                    for (uint32_t j = lane; j < p.k; j += 32) out[j] = sel[i].buf[j];
                }
            }
        }
    }
}

static inline size_t list_kernel_smem(uint32_t row_bytes, uint32_t q_bytes, uint32_t nsort) {
    return static_cast<size_t>(CTA_QUERIES) * q_bytes + CTA_QUERIES * 8 + CTA_QUERIES * 8 + 2ull * TILE_ROWS * smem_row_pitch(row_bytes) +
           static_cast<size_t>(CTA_QUERIES) * nsort * 8;
}

// ---------------------------------------------------------------------------
// Probe selection: get_centroids_dist + select_probed_clusters
// (src/utils/k_means_utils.rs:76-99, 3007-3029).  One CTA per query sorts all
// (distance, cell) pairs -- equal distances in ascending cell id -- and walks the
// ranking until `nprobe` cells are chosen AND at least k vectors are reachable
// (empty cells count toward nprobe only).  List sizes are global, so every shard
// of a sharded index derives the same probe list.
// ---------------------------------------------------------------------------
struct ProbeParams {
    const float* cdist;      // [nq][nlist]
    uint64_t nq;
    uint32_t nlist, nlist_pow2;
    const uint64_t* offsets; // [nlist+1]
    uint32_t nprobe;
    uint64_t k;
    uint32_t* probes;        // [nq][probe_pitch]
    uint32_t probe_pitch;
    uint32_t* n_probes;      // [nq]
    uint32_t* overflow;      // set to 1 if some query needs more than probe_pitch cells
    unsigned long long* stat_scanned;  // sum over queries of reachable vectors
    unsigned long long* stat_probed;   // sum over queries of probed cells
    uint32_t list_begin, list_end;     // this shard's lists: stat_scanned[1] counts only their vectors
};

__global__ void __launch_bounds__(256) probe_kernel(ProbeParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem);
    __shared__ uint32_t s_np;
    const uint64_t q = blockIdx.x;
    const float* row = p.cdist + q * p.nlist;
    for (uint32_t i = threadIdx.x; i < p.nlist_pow2; i += blockDim.x) keys[i] = (i < p.nlist) ? make_key(row[i], i) : KEY_SENTINEL;
    __syncthreads();
    bitonic_sort_keys<true>(keys, p.nlist_pow2, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) {
        uint32_t chosen = 0;
        uint64_t reach = 0, local = 0;
        for (uint32_t i = 0; i < p.nlist; i++) {
            uint32_t c = key_idx(keys[i]);
            chosen++;
            reach += p.offsets[c + 1] - p.offsets[c];
            if (c >= p.list_begin && c < p.list_end) local += p.offsets[c + 1] - p.offsets[c];
            if (chosen >= p.nprobe && reach >= p.k) break;
        }
        s_np = chosen;
        if (chosen > p.probe_pitch) atomicExch(p.overflow, 1u);
        p.n_probes[q] = min(chosen, p.probe_pitch);
        atomicAdd(p.stat_scanned, static_cast<unsigned long long>(reach));
        atomicAdd(p.stat_probed, static_cast<unsigned long long>(chosen));
        atomicAdd(p.stat_probed + 1, static_cast<unsigned long long>(local));
    }
    __syncthreads();
    const uint32_t np = min(s_np, p.probe_pitch);
    for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) p.probes[q * p.probe_pitch + i] = key_idx(keys[i]);
}

// Probe-expansion walk over an already ranked prefix: `ranked` holds, per query, the `pitch` nearest (distance, cell)
// keys in ascending order (tile_kernel<EPI_SELECT> over the centroid table).  One thread per query applies the
// select_probed_clusters rule; if the prefix is exhausted before the rule is satisfied the query is flagged and the
// caller repeats the batch with the full ranking (probe_kernel).
__global__ void probe_walk_kernel(const uint64_t* __restrict__ ranked, ProbeParams p) {
    const uint64_t q = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (q >= p.nq) return;
    const uint64_t* keys = ranked + q * p.probe_pitch;
    uint32_t chosen = 0;
    uint64_t reach = 0, local = 0;
    bool done = false;
    for (uint32_t i = 0; i < p.probe_pitch; i++) {
        const uint32_t c = key_idx(keys[i]);
        if (c == IDX_INVALID) break;
        p.probes[q * p.probe_pitch + i] = c;
        chosen++;
        reach += p.offsets[c + 1] - p.offsets[c];
        if (c >= p.list_begin && c < p.list_end) local += p.offsets[c + 1] - p.offsets[c];
        if (chosen >= p.nprobe && reach >= p.k) { done = true; break; }
    }
    if (!done && chosen < p.nlist) atomicExch(p.overflow, 1u);
    p.n_probes[q] = chosen;
    atomicAdd(p.stat_scanned, static_cast<unsigned long long>(reach));
    atomicAdd(p.stat_probed, static_cast<unsigned long long>(chosen));
    atomicAdd(p.stat_probed + 1, static_cast<unsigned long long>(local));
}

// ---------------------------------------------------------------------------
// Finalize: merge `parts` partial key lists per query, emit the k best.
//   id = id_map ? id_map[idx] : idx + id_base ; output row = row_map ? row_map[q] : q
// ---------------------------------------------------------------------------
struct FinalizeParams {
    const uint64_t* part_keys;  // [nq][parts][kc]
    uint32_t parts, kc, k, nsort;
    uint64_t nq;
    const uint64_t* id_map;
    uint64_t id_base;
    const uint64_t* row_map;
    const uint32_t* parts_used;  // optional [nq]: only the first parts_used[q] lists of a query hold data
    uint64_t* out_ids;
    float* out_dist;
    uint32_t* out_counts;
};

__global__ void __launch_bounds__(512) finalize_kernel(FinalizeParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem);
    const uint64_t q = blockIdx.x;
    const uint32_t parts_q = p.parts_used ? min(p.parts, p.parts_used[q]) : p.parts;
    const uint32_t total = parts_q * p.kc;
    const uint32_t nsort = min(p.nsort, next_pow2(max(max(total, p.k), 2u)));   // sort only what this query filled
    const uint64_t* src = p.part_keys + q * (static_cast<uint64_t>(p.parts) * p.kc);
    for (uint32_t i = threadIdx.x; i < nsort; i += blockDim.x) keys[i] = (i < total) ? src[i] : KEY_SENTINEL;
    __syncthreads();
    bitonic_sort_keys<true>(keys, nsort, threadIdx.x, blockDim.x);
    const uint64_t orow = p.row_map ? p.row_map[q] : q;
    uint32_t my_valid = 0;
    for (uint32_t j = threadIdx.x; j < p.k; j += blockDim.x) {
        uint64_t key = keys[j];
        bool ok = key_idx(key) != IDX_INVALID;
        uint64_t id = 0xFFFFFFFFFFFFFFFFull;
        float d = INFINITY;
        if (ok) {
            uint32_t idx = key_idx(key);
            id = p.id_map ? p.id_map[idx] : (static_cast<uint64_t>(idx) + p.id_base);
            d = key_dist(key);
            my_valid++;
        }
        p.out_ids[orow * p.k + j] = id;
        if (p.out_dist) p.out_dist[orow * p.k + j] = d;
    }
    if (p.out_counts) {
        __shared__ uint32_t s_cnt;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        if (my_valid) atomicAdd(&s_cnt, my_valid);
        __syncthreads();
        if (threadIdx.x == 0) p.out_counts[orow] = s_cnt;
    }
}

// Cross-shard merge (annb_merge_topk_dev): inputs are final (id, dist) rows, [part][nq][k].
// Order: (distance, id).
struct MergeParams {
    const uint64_t* part_ids;
    const float* part_dist;
    uint32_t parts, k, nsort;
    uint64_t nq;
    uint64_t* out_ids;
    float* out_dist;
    uint32_t* out_counts;
};

__global__ void __launch_bounds__(128) merge_kernel(MergeParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    // sort (ordered dist, slot) keys, then look ids up by slot: ids are 64-bit and do not fit the key
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem);
    const uint64_t q = blockIdx.x;
    const uint32_t total = p.parts * p.k;
    // Ties on distance must break on id, not slot: use a two-pass approach -- sort by (dist, id_lo32)
    // when ids fit 32 bits is not general, so keys carry the slot and ties are fixed up below.
    for (uint32_t i = threadIdx.x; i < p.nsort; i += blockDim.x) {
        uint64_t key = KEY_SENTINEL;
        if (i < total) {
            uint32_t part = i / p.k, j = i - part * p.k;
            uint64_t src = (static_cast<uint64_t>(part) * p.nq + q) * p.k + j;
            if (p.part_ids[src] != 0xFFFFFFFFFFFFFFFFull) key = make_key(p.part_dist[src], i);
        }
        keys[i] = key;
    }
    __syncthreads();
    bitonic_sort_keys<true>(keys, p.nsort, threadIdx.x, blockDim.x);
    // equal-distance runs: order by id (serial insertion sort inside each run; runs are short)
    if (threadIdx.x == 0) {
        uint32_t i = 0;
        while (i < total && key_idx(keys[i]) != IDX_INVALID) {
            uint32_t j = i + 1;
            while (j < total && key_idx(keys[j]) != IDX_INVALID && (keys[j] >> 32) == (keys[i] >> 32)) j++;
            for (uint32_t a = i + 1; a < j; a++) {
                uint64_t ka = keys[a];
                uint32_t sa = key_idx(ka);
                uint64_t ida = p.part_ids[(static_cast<uint64_t>(sa / p.k) * p.nq + q) * p.k + (sa % p.k)];
                uint32_t b = a;
                while (b > i) {
                    uint32_t sb = key_idx(keys[b - 1]);
                    uint64_t idb = p.part_ids[(static_cast<uint64_t>(sb / p.k) * p.nq + q) * p.k + (sb % p.k)];
                    if (idb <= ida) break;
                    keys[b] = keys[b - 1];
                    b--;
                }
                keys[b] = ka;
            }
            i = j;
        }
    }
    __syncthreads();
    uint32_t my_valid = 0;
    for (uint32_t j = threadIdx.x; j < p.k; j += blockDim.x) {
        uint64_t key = keys[j];
        uint64_t id = 0xFFFFFFFFFFFFFFFFull;
        float d = INFINITY;
        if (key_idx(key) != IDX_INVALID) {
            uint32_t s = key_idx(key);
            uint64_t src = (static_cast<uint64_t>(s / p.k) * p.nq + q) * p.k + (s % p.k);
            id = p.part_ids[src];
            d = p.part_dist[src];
            my_valid++;
        }
        p.out_ids[q * p.k + j] = id;
        if (p.out_dist) p.out_dist[q * p.k + j] = d;
    }
    if (p.out_counts) {
        __shared__ uint32_t s_cnt;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        if (my_valid) atomicAdd(&s_cnt, my_valid);
        __syncthreads();
        if (threadIdx.x == 0) p.out_counts[q] = s_cnt;
    }
}

// Cross-shard merge of interleaved per-shard results (annb_merge_shards_dev): shard p's [nq][k] ids start at
// base + p * part_stride, its [nq][k] distances dist_offset bytes further on -- one buffer per shard, so one all-gather
// (or one peer copy per shard) moves ids and distances together.  Order: (distance, shard, position inside the shard's
// own list).  Every shard's list is already in the index's own order -- (distance, row) for flat shards, (distance, list
// position) for IVF shards -- and shards own ascending, disjoint row / list ranges, so with the shards passed in that
// order this IS the order of the unsharded index (src/utils/heap_structs.rs:12-38, src/cpu/ivf.rs:367-381).
struct MergeShardsParams {
    const uint8_t* base;
    uint64_t part_stride, dist_offset;
    uint32_t parts, k;
    uint64_t nq;
    uint64_t* out_ids;
    float* out_dist;
    uint32_t* out_counts;
    // fused shard check (merge_shards_sort_kernel<true>): every shard's bounds [nq] sit bound_offset bytes into its block, its int32
    // status word behind them; verdict[0] counts my_part's queries to refine, verdict[1]: bit 0 some shard has to refine, bit 1 some
    // shard's call failed
    uint64_t bound_offset;
    uint32_t my_part;
    uint32_t* verdict;
};

// Small merges (parts * k <= 128, e.g. 8 shards x k = 10 or 15): one warp per query sorts the (distance, slot) keys of all shards
// in registers -- slot = shard * k + position, so the key order IS (distance, shard, position) -- and emits the first k.  The
// binary-search variant below walks ~50 dependent loads per entry; this one is one load round trip and ~500 register ops.
// CHECK: the merged verdict of annb_shard_check_gathered_dev in the same pass -- the k-th merged key is in a register, every
// shard's bound for the query is one more load: no second kernel over the merged rows.
template <bool CHECK>
__global__ void __launch_bounds__(128) merge_shards_sort_kernel(MergeShardsParams p) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t q = static_cast<uint64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= p.nq) return;
    const uint32_t total = p.parts * p.k;
    uint64_t key[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t i = lane + 32u * j;
        key[j] = KEY_SENTINEL;
        if (i < total) {
            const uint32_t part = i / p.k, pos = i - part * p.k;
            const uint8_t* pb = p.base + static_cast<uint64_t>(part) * p.part_stride;
            if (reinterpret_cast<const uint64_t*>(pb)[q * p.k + pos] != 0xFFFFFFFFFFFFFFFFull)
                key[j] = make_key(reinterpret_cast<const float*>(pb + p.dist_offset)[q * p.k + pos], i);
        }
    }
    warp_bitonic_sort_regs<4>(key, lane);
    uint32_t valid = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t r = lane + 32u * j;            // rank of key[j]
        if (r < p.k) {
            uint64_t id = 0xFFFFFFFFFFFFFFFFull;
            float d = INFINITY;
            if (key[j] != KEY_SENTINEL) {
                const uint32_t i = key_idx(key[j]), part = i / p.k, pos = i - part * p.k;
                const uint8_t* pb = p.base + static_cast<uint64_t>(part) * p.part_stride;
                id = reinterpret_cast<const uint64_t*>(pb)[q * p.k + pos];
                d = reinterpret_cast<const float*>(pb + p.dist_offset)[q * p.k + pos];
                valid++;
            }
            p.out_ids[q * p.k + r] = id;
            if (p.out_dist) p.out_dist[q * p.k + r] = d;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) valid += __shfl_xor_sync(0xFFFFFFFFu, valid, off);
    if (p.out_counts && lane == 0) p.out_counts[q] = valid;
    if constexpr (CHECK) {
        // k-th merged distance (+inf where fewer than k rows were found at all): rank k - 1 lives in key[(k - 1) / 32] of lane (k - 1) % 32
        const uint32_t kr = p.k - 1;
        uint64_t kk = KEY_SENTINEL;
#pragma unroll
        for (int j = 0; j < 4; j++) if (static_cast<uint32_t>(j) == (kr >> 5)) kk = key[j];
        kk = __shfl_sync(0xFFFFFFFFu, kk, kr & 31u);
        const float dk = kk != KEY_SENTINEL ? key_dist(kk) : INFINITY;
        uint32_t flags = 0, mine = 0;
        for (uint32_t s = lane; s < p.parts; s += 32) {
            const uint8_t* bb = p.base + static_cast<uint64_t>(s) * p.part_stride + p.bound_offset;
            const float b = reinterpret_cast<const float*>(bb)[q];
            const bool need = !(b > dk) && b != INFINITY;
            if (need) { flags |= 1u; if (s == p.my_part) mine = 1; }
            if (q == 0 && *reinterpret_cast<const int32_t*>(bb + p.nq * 4) != 0) flags |= 2u;   // a failed shard poisons the step on every rank alike
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { flags |= __shfl_xor_sync(0xFFFFFFFFu, flags, off); mine |= __shfl_xor_sync(0xFFFFFFFFu, mine, off); }
        if (lane == 0) {
            if (mine) atomicAdd(p.verdict, 1u);
            if (flags) atomicOr(p.verdict + 1, flags);
        }
    }
}

// One warp per query.  Each per-shard list is ascending, so an entry's final rank is its own slot plus, for every other
// shard, the number of that shard's entries that precede it (<= for earlier shards, < for later ones): k * parts * parts
// compares per query, no sort, no shared memory.
__global__ void __launch_bounds__(128) merge_shards_kernel(MergeShardsParams p) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t q = static_cast<uint64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= p.nq) return;
    const uint32_t total = p.parts * p.k;
    uint32_t valid = 0;
    for (uint32_t i = lane; i < total; i += 32) {
        const uint32_t part = i / p.k, j = i - part * p.k;
        const uint8_t* pb = p.base + static_cast<uint64_t>(part) * p.part_stride;
        const uint64_t id = reinterpret_cast<const uint64_t*>(pb)[q * p.k + j];
        if (id == 0xFFFFFFFFFFFFFFFFull) continue;
        const float d = reinterpret_cast<const float*>(pb + p.dist_offset)[q * p.k + j];
        const uint32_t od = f32_to_ordered(d);
        uint32_t rank = j;
        for (uint32_t o = 0; o < p.parts && rank < p.k; o++) {
            if (o == part) continue;
            const uint8_t* ob = p.base + static_cast<uint64_t>(o) * p.part_stride;
            const uint64_t* oids = reinterpret_cast<const uint64_t*>(ob) + q * p.k;
            const float* odist = reinterpret_cast<const float*>(ob + p.dist_offset) + q * p.k;
            // number of entries of shard o that come before (d, part): binary search over its ascending list
            uint32_t lo = 0, hi = p.k;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                const bool before = oids[mid] != 0xFFFFFFFFFFFFFFFFull &&
                                    (o < part ? f32_to_ordered(odist[mid]) <= od : f32_to_ordered(odist[mid]) < od);
                if (before) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < p.k) {
            p.out_ids[q * p.k + rank] = id;
            if (p.out_dist) p.out_dist[q * p.k + rank] = d;
            valid++;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) valid += __shfl_xor_sync(0xFFFFFFFFu, valid, off);
    // fewer than k valid entries in total: pad the tail
    for (uint32_t j = valid + lane; j < p.k; j += 32) {
        p.out_ids[q * p.k + j] = 0xFFFFFFFFFFFFFFFFull;
        if (p.out_dist) p.out_dist[q * p.k + j] = INFINITY;
    }
    if (p.out_counts && lane == 0) p.out_counts[q] = valid;
}

}  // namespace annb
