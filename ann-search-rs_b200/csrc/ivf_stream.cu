// ivf_stream.cu -- the HBM-streaming IVF list scan (query-major): the kernel of small query batches and of the exact
// fallback, and the one the north-star's "list scan at HBM speed" bound is about (src/cpu/ivf.rs:367-381, one probed list
// after the other, rows contiguous in list order).
//
// One warp = one (query, part [, list segment]) and walks its probed lists in 32-row tiles.  The tiles do not go through
// per-lane cp.async: lane 0 of every warp drives a private ring of TMA tile loads (cp.async.bulk.tensor.2d, one 128-byte
// K slab x 32 rows per load, SWIZZLE_128B, completion on an mbarrier), three or four tiles deep: two tiles per warp are in
// flight while one is scored, and the SM's resident warps (4 for f32 rows of 128, 8 for bf16, 12 for SQ8 -- the ring is the
// only large shared-memory user) together keep well over the HBM latency-bandwidth product in flight.  The 128-byte swizzle makes the
// lane-per-row reads conflict-free (lane r reads 16-byte chunk j of its row at position j ^ (r & 7) of the slab row), and
// TMA zero-fills rows past the end of the index.  Rows of the box that belong to the next list are loaded and ignored.
// Arithmetic: refdist.cuh's reference order (8 lane accumulators over consecutive chunks, the reference's reduce tree,
// scalar tail; fused multiply-add for bf16 rows, exact integers for SQ8), dequantisation fused into the walk -- results are
// bit-identical to the CPU path and to ivf_scan_kernel (simt_kernels.cuh), which stays as the fallback for handles without
// a tensor map.
#include <cuda.h>

#include <algorithm>

#include "flat_tc.hpp"
#include "index.hpp"
#include "tc_common.cuh"

namespace annb {

int tc_make_tmap(CUtensorMap* tm, void* base, uint64_t rows, uint32_t kp_elems, int elem_bytes, uint32_t box_rows);

namespace tc {

constexpr int STREAM_WARPS = 4;
constexpr int STREAM_ROWS = 32;                       // rows per tile = one row per lane
constexpr int STREAM_SLAB = STREAM_ROWS * SLAB_BYTES; // 4 KiB: 32 rows x 128 B

struct StreamParams {
    const float* row_norms;
    const int32_t* row_norms_i;
    const uint8_t* queries;      // scan-side queries [nq][q_bytes]
    uint32_t q_bytes, row_bytes;
    uint64_t nq;
    uint32_t dim;
    int bf16_self;
    const uint32_t* probes;      // [nq][probe_pitch] cell ids in rank order
    uint32_t probe_pitch;
    const uint32_t* n_probes;
    const uint64_t* offsets;     // global CSR offsets
    uint32_t list_begin, list_end;
    uint64_t shard_row0;
    uint32_t parts, subs, k, nsort;
    uint32_t nslab, n_stages;    // K slabs per row, ring depth (tiles)
    uint64_t* part_keys;         // [nq][parts * subs][k]
};

// 16-byte chunk g of row r of a staged tile ([nslab][32 rows][128 B], 128-byte swizzle)
__device__ __forceinline__ const uint8_t* swz(const uint8_t* tile, uint32_t r, uint32_t g) {
    return tile + (g >> 3) * STREAM_SLAB + r * SLAB_BYTES + (((g & 7u) ^ (r & 7u)) << 4);
}

template <int ELEM>
__device__ __forceinline__ void load8_swz(const uint8_t* tile, uint32_t r, uint32_t c, float v[8]) {
    if (ELEM == 4) {
        const float4 a = *reinterpret_cast<const float4*>(swz(tile, r, 2 * c));
        const float4 b = *reinterpret_cast<const float4*>(swz(tile, r, 2 * c + 1));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        const uint4 a = *reinterpret_cast<const uint4*>(swz(tile, r, c));
        v[0] = bf16_bits_to_f32(a.x & 0xFFFFu); v[1] = bf16_bits_to_f32(a.x >> 16);
        v[2] = bf16_bits_to_f32(a.y & 0xFFFFu); v[3] = bf16_bits_to_f32(a.y >> 16);
        v[4] = bf16_bits_to_f32(a.z & 0xFFFFu); v[5] = bf16_bits_to_f32(a.z >> 16);
        v[6] = bf16_bits_to_f32(a.w & 0xFFFFu); v[7] = bf16_bits_to_f32(a.w >> 16);
    }
}
template <int ELEM>
__device__ __forceinline__ float load1_swz(const uint8_t* tile, uint32_t r, uint32_t e) {
    const uint32_t byte = e * ELEM;
    const uint8_t* p = swz(tile, r, byte >> 4) + (byte & 15u);
    if (ELEM == 4) return *reinterpret_cast<const float*>(p);
    return bf16_bits_to_f32(*reinterpret_cast<const uint16_t*>(p));
}

// RT: 0 f32 rows, 1 bf16 rows, 2 int8 codes.  QT: query element type.
template <int RT, int QT, int MET>
__global__ void __launch_bounds__(STREAM_WARPS * 32) ivf_stream_kernel(const __grid_constant__ CUtensorMap tm_rows, const StreamParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    constexpr int RELEM = (RT == 0) ? 4 : (RT == 1 ? 2 : 1);
    constexpr int QELEM = (QT == QT_F32) ? 4 : (QT == QT_BF16 ? 2 : 1);
    const uint32_t tile_bytes = p.nslab * STREAM_SLAB;
    const uint32_t ring_bytes = p.n_stages * tile_bytes;                       // per warp, a multiple of 4 KiB
    uint8_t* ring = smem + warp * ring_bytes;
    uint8_t* tail = smem + STREAM_WARPS * ring_bytes + warp * (8 * 8 + p.q_bytes + p.nsort * 8u);   // per warp: barriers | query | select buffer
    uint64_t* bar = reinterpret_cast<uint64_t*>(tail);                         // [<= 8]
    uint64_t* s_sel = reinterpret_cast<uint64_t*>(tail + 64);                  // [nsort]
    uint8_t* s_q = tail + 64 + p.nsort * 8u;                                   // [q_bytes]

    const uint64_t task = static_cast<uint64_t>(blockIdx.x) * STREAM_WARPS + warp;
    const uint32_t slices = p.parts * p.subs;
    const uint64_t q = task / slices;
    const uint32_t slice = static_cast<uint32_t>(task % slices);
    const uint32_t part = slice / p.subs, sub = slice - part * p.subs;
    if (q >= p.nq) return;  // whole warp exits together (task is warp-uniform); warps never meet at a block barrier

    if (lane == 0) {
        for (uint32_t s = 0; s < p.n_stages; s++) mbar_init(bar + s, 1);
        fence_barrier_init();
        fence_proxy_async();
        tma_prefetch_desc(&tm_rows);
    }
    for (uint32_t c = lane; c < (p.q_bytes >> 4); c += 32)
        *reinterpret_cast<uint4*>(s_q + (c << 4)) = *reinterpret_cast<const uint4*>(p.queries + q * p.q_bytes + (c << 4));
    __syncwarp();
    float qnorm = 1.0f;
    int32_t qnsq = 0;
    if (QT == QT_I8) {
        for (uint32_t e = lane; e < p.dim; e += 32) {
            const int32_t v = reinterpret_cast<const int8_t*>(s_q)[e];
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

    // tile generator over this part's probe ranks (part, part + parts, ...); every lane runs it identically
    uint32_t rank = part;
    uint64_t pos = 0, end = 0;  // current list: shard-local row range [pos, end)
    auto next_tile = [&](uint64_t& t_row, uint32_t& t_n) -> bool {
        while (pos >= end) {
            if (rank >= np) return false;
            const uint32_t c = probes[rank];
            rank += p.parts;
            if (c < p.list_begin || c >= p.list_end) continue;  // list lives on another shard
            pos = p.offsets[c] - p.shard_row0;
            end = p.offsets[c + 1] - p.shard_row0;
            if (p.subs > 1) {   // this warp's segment of the list (whole tiles)
                const uint64_t seg = ((end - pos + p.subs - 1) / p.subs + STREAM_ROWS - 1) / STREAM_ROWS * STREAM_ROWS;
                pos = min(end, pos + sub * seg);
                end = min(end, pos + seg);
            }
        }
        t_row = pos;
        t_n = static_cast<uint32_t>(min(static_cast<uint64_t>(STREAM_ROWS), end - pos));
        pos += t_n;
        return true;
    };
    auto issue = [&](uint32_t stage, uint64_t row) {   // lane 0: one TMA load per K slab of the tile
        mbar_expect_tx(bar + stage, tile_bytes);
        for (uint32_t s = 0; s < p.nslab; s++)
            tma_load_2d(smem_u32(ring + stage * tile_bytes + s * STREAM_SLAB), &tm_rows, bar + stage, static_cast<int32_t>(s * (SLAB_BYTES / RELEM)),
                        static_cast<int32_t>(row));
    };

    // prologue: fill the ring.  The generator runs n_stages tiles ahead of the consumer; (row, n) of the tiles in flight
    // live in a small per-lane queue (registers, identical in every lane).
    uint64_t q_row[8];
    uint32_t q_n[8];
    uint32_t produced = 0;
    bool more = true;
#pragma unroll
    for (int s = 0; s < 8; s++) {
        q_row[s] = 0; q_n[s] = 0;
        if (static_cast<uint32_t>(s) < p.n_stages && more) {
            more = next_tile(q_row[s], q_n[s]);
            if (more) {
                if (lane == 0) issue(s, q_row[s]);
                produced++;
            }
        }
    }
    uint32_t consumed = 0;
    while (consumed < produced) {
        const uint32_t stage = consumed % p.n_stages, ph = (consumed / p.n_stages) & 1u;
        uint64_t cur_row = 0;
        uint32_t cur_n = 0;
#pragma unroll
        for (int s = 0; s < 8; s++)
            if (static_cast<uint32_t>(s) == stage) { cur_row = q_row[s]; cur_n = q_n[s]; }
        mbar_wait(bar + stage, ph);
        const uint8_t* tile = ring + stage * tile_bytes;
        const bool valid = lane < cur_n;
        const uint64_t row = cur_row + lane;
        float dist = 0.0f;
        if (RT == 2) {
            // SQ8: exact code-space integers (src/utils/dist.rs:5015-5077), whole 16-code chunks through dp4a
            int32_t dot = 0, xx = 0;
            const uint32_t chunks = (p.dim + 15u) >> 4;
#pragma unroll 4
            for (uint32_t c = 0; c < chunks; c++) {
                const int4 x = *reinterpret_cast<const int4*>(swz(tile, lane, c));
                const int4 y = *reinterpret_cast<const int4*>(s_q + c * 16);
                xx = __dp4a(x.x, x.x, xx); xx = __dp4a(x.y, x.y, xx); xx = __dp4a(x.z, x.z, xx); xx = __dp4a(x.w, x.w, xx);
                dot = __dp4a(x.x, y.x, dot); dot = __dp4a(x.y, y.y, dot); dot = __dp4a(x.z, y.z, dot); dot = __dp4a(x.w, y.w, dot);
            }
            const int32_t xn = (MET == MET_COS && valid) ? p.row_norms_i[row] : 0;
            dist = finish_i8<MET>(dot, xx, qnsq, xn);
        } else {
            constexpr bool FMA = (RELEM == 2);
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = 0.0f;
            const uint32_t chunks = p.dim >> 3;
            for (uint32_t c = 0; c < chunks; c++) {     // (not unrolled: unrolling by 4 cost the f32 kernel a quarter of its rate)
                float x[8], y[8];
                load8_swz<RELEM>(tile, lane, c, x);
                load8<QELEM>(s_q + c * 8 * QELEM, y);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (MET == MET_L2) {
                        const float d = __fsub_rn(x[j], y[j]);
                        acc[j] = FMA ? __fmaf_rn(d, d, acc[j]) : __fadd_rn(acc[j], __fmul_rn(d, d));
                    } else {
                        acc[j] = FMA ? __fmaf_rn(x[j], y[j], acc[j]) : __fadd_rn(acc[j], __fmul_rn(x[j], y[j]));
                    }
                }
            }
            float sum = FMA ? hsum_bf16path(acc) : hsum_wide(acc);
            for (uint32_t e = chunks * 8; e < p.dim; e++) {  // scalar tail: `sum += d * d` (not fused)
                const float xv = load1_swz<RELEM>(tile, lane, e);
                const float yv = load1<QELEM>(s_q, e);
                if (MET == MET_L2) {
                    const float d = __fsub_rn(xv, yv);
                    sum = __fadd_rn(sum, __fmul_rn(d, d));
                } else {
                    sum = __fadd_rn(sum, __fmul_rn(xv, yv));
                }
            }
            const float xn = (MET == MET_COS && valid) ? p.row_norms[row] : 1.0f;
            dist = finish_fp<MET>(sum, qnorm, xn);
        }
        __syncwarp();                       // every lane has read the tile: the stage can be refilled
        if (more) {
            uint64_t n_row = 0;
            uint32_t n_n = 0;
            more = next_tile(n_row, n_n);
            if (more) {
#pragma unroll
                for (int s = 0; s < 8; s++)
                    if (static_cast<uint32_t>(s) == stage) { q_row[s] = n_row; q_n[s] = n_n; }
                if (lane == 0) {
                    fence_proxy_async();    // generic-proxy reads of the stage before the async-proxy refill
                    issue(stage, n_row);
                }
                produced++;
            }
        }
        sel.offer(make_key(dist, static_cast<uint32_t>(row)), valid);
        consumed++;
    }
    sel.flush();
    uint64_t* out = p.part_keys + (q * slices + slice) * p.k;
    for (uint32_t j = lane; j < p.k; j += 32) out[j] = sel.buf[j];
}

}  // namespace tc

struct StreamState {
    CUtensorMap tm;
    uint32_t nslab = 0;
};

int tc_stream_prepare(annb_index* ix) {
    if (!ix->is_ivf || ix->n == 0) return ANNB_OK;
    const uint32_t elem = elem_bytes(ix->dtype);
    if (ix->row_bytes > 2048) return ANNB_OK;                 // 16 K slabs per row at most: wider rows keep the cp.async kernel
    StreamState* st = new StreamState();
    st->nslab = ceil_div(ix->row_bytes, static_cast<uint32_t>(tc::SLAB_BYTES));
    int rc = tc_make_tmap(&st->tm, ix->d_rows, ix->n, ix->row_bytes / elem, static_cast<int>(elem), tc::STREAM_ROWS);
    if (rc != ANNB_OK) { delete st; return ANNB_OK; }         // no tensor map (old driver): the cp.async kernel serves the index
    ix->tc_stream = st;
    return ANNB_OK;
}
void tc_stream_destroy(annb_index* ix) {
    delete ix->tc_stream;
    ix->tc_stream = nullptr;
}
bool tc_stream_supported(const annb_index* ix) { return ix->tc_stream != nullptr && ix->opt_ivf_stream != 0; }

template <int RT, int QT, int MET>
static int launch_stream(const CUtensorMap& tm, const tc::StreamParams& p, uint32_t grid, size_t smem, cudaStream_t s) {
    auto kern = tc::ivf_stream_kernel<RT, QT, MET>;
    ANNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, tc::STREAM_WARPS * 32, smem, s>>>(tm, p);
    ANNB_CUDA_CHECK(cudaGetLastError());
    return ANNB_OK;
}

int tc_stream_scan(annb_index* ix, const StreamScanArgs& a, cudaStream_t s) {
    StreamState* st = ix->tc_stream;
    tc::StreamParams p{};
    p.row_norms = ix->d_norms; p.row_norms_i = ix->d_norms_i; p.queries = a.queries; p.q_bytes = a.q_bytes; p.row_bytes = ix->row_bytes;
    p.nq = a.nq; p.dim = ix->dim; p.bf16_self = a.bf16_self; p.probes = a.probes; p.probe_pitch = a.probe_pitch; p.n_probes = a.n_probes;
    p.offsets = ix->d_offsets; p.list_begin = ix->list_begin; p.list_end = ix->list_end; p.shard_row0 = ix->shard_row0;
    p.parts = a.parts; p.subs = a.subs; p.k = a.k; p.nsort = a.nsort; p.part_keys = a.part_keys;
    p.nslab = st->nslab;
    const size_t tile = static_cast<size_t>(st->nslab) * tc::STREAM_SLAB;
    const size_t per_warp_tail = 64 + a.q_bytes + static_cast<size_t>(a.nsort) * 8;
    // ring depth: three tiles per warp (four for the 4 KiB tiles of narrow rows).  Bytes in flight come from the number of
    // resident warps, not from deep rings: f32 d = 128 -> 48 KiB per warp, one CTA per SM (already HBM-bound); bf16 -> 24 KiB,
    // two CTAs; SQ8 -> 16 KiB, three CTAs -- the narrower the rows, the more warps hide the per-tile latencies.
    const size_t budget = 220 * 1024;
    size_t stages = tile <= 4096 ? 4 : 3;
    while (stages > 2 && tc::STREAM_WARPS * (stages * tile + per_warp_tail) > budget) stages--;
    if (stages < 2 || tc::STREAM_WARPS * (stages * tile + per_warp_tail) > budget) { set_last_error("streaming scan: rows too wide for the tile ring"); return ANNB_ERR_UNSUPPORTED; }
    p.n_stages = static_cast<uint32_t>(stages);
    const size_t smem = tc::STREAM_WARPS * (stages * tile + per_warp_tail);
    const uint32_t grid = static_cast<uint32_t>(ceil_div<uint64_t>(a.nq * a.parts * a.subs, tc::STREAM_WARPS));
    const bool cos = ix->metric == ANNB_COSINE;
    const int rt = ix->dtype, qt = a.qt;
    if (rt == ANNB_F32 && qt == QT_F32) return cos ? launch_stream<0, QT_F32, MET_COS>(st->tm, p, grid, smem, s) : launch_stream<0, QT_F32, MET_L2>(st->tm, p, grid, smem, s);
    if (rt == ANNB_BF16 && qt == QT_F32) return cos ? launch_stream<1, QT_F32, MET_COS>(st->tm, p, grid, smem, s) : launch_stream<1, QT_F32, MET_L2>(st->tm, p, grid, smem, s);
    if (rt == ANNB_BF16 && qt == QT_BF16) return cos ? launch_stream<1, QT_BF16, MET_COS>(st->tm, p, grid, smem, s) : launch_stream<1, QT_BF16, MET_L2>(st->tm, p, grid, smem, s);
    if (rt == ANNB_SQ8 && qt == QT_I8) return cos ? launch_stream<2, QT_I8, MET_COS>(st->tm, p, grid, smem, s) : launch_stream<2, QT_I8, MET_L2>(st->tm, p, grid, smem, s);
    set_last_error("unsupported (row, query) type pair");
    return ANNB_ERR_INVALID_ARGUMENT;
}

}  // namespace annb
