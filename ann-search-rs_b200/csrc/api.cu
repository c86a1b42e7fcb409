// api.cu -- the C ABI (include/annb200.h) and the host-side orchestration of the kernels.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "index.hpp"
#include "prep_kernels.cuh"
#include "simt_kernels.cuh"
#include "flat_tc.hpp"

namespace annb {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

static int fail(int code, const std::string& msg) {
    set_last_error(msg);
    return code;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { (void)cudaGetLastError(); return; }
        ok = (cudaSetDevice(dev) == cudaSuccess);
        if (!ok) (void)cudaGetLastError();
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

#define ANNB_DEVICE(dev)                                                                        \
    ::annb::DeviceGuard _guard(dev);                                                            \
    if (!_guard.ok) return ::annb::fail(ANNB_ERR_CUDA, "no usable CUDA device " + std::to_string(dev) + \
                                        " (libannb200 has no CPU fallback)")

// Dominant-kernel timing: bracket a launch with events on its own stream (option "time_kernels").
struct KernelTimer {
    annb_index* ix;
    cudaStream_t s;
    cudaEvent_t a = nullptr, b = nullptr;
    KernelTimer(annb_index* ix_, cudaStream_t s_, bool enable = true) : ix(ix_), s(s_) {
        if (!ix->opt_time_kernels || !enable) return;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { a = b = nullptr; (void)cudaGetLastError(); return; }
        cudaEventRecord(a, s);
    }
    ~KernelTimer() {
        if (!a) return;
        cudaEventRecord(b, s);
        ix->timed.emplace_back(a, b);
    }
};
static void collect_timers(const annb_index* ix) {
    for (auto& pr : ix->timed) {
        float ms = 0.f;
        if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
            ix->timed_ms_total += ms;
            ix->timed_launches++;
        }
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    (void)cudaGetLastError();
    ix->timed.clear();
}

static inline uint32_t grid_for(uint64_t work, uint32_t block, uint32_t cap = 148 * 32) {
    uint64_t g = (work + block - 1) / block;
    return static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(g, cap)));
}

template <typename T>
static int dmalloc(T** p, size_t count, annb_index* ix) {
    size_t bytes = std::max<size_t>(count * sizeof(T), 16);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), bytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? ANNB_ERR_OUT_OF_MEMORY : ANNB_ERR_CUDA,
                    std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    }
    if (ix) ix->device_bytes += bytes;
    return ANNB_OK;
}

// ---------------------------------------------------------------------------
// kernel dispatch
// ---------------------------------------------------------------------------
template <int RT, int QT, int MET, int EPI>
static int launch_tile(const TileParams& p, dim3 grid, size_t smem, cudaStream_t s) {
    auto kern = tile_kernel<RT, QT, MET, EPI>;
    if (smem > 227 * 1024) return fail(ANNB_ERR_UNSUPPORTED, "tile_kernel: dim/k too large for shared memory (DimTooHighForSharedMemory)");
    ANNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, TILE_THREADS, smem, s>>>(p);
    ANNB_CUDA_CHECK(cudaGetLastError());
    return ANNB_OK;
}

// rt: index dtype; qt: query element type
static int launch_select_tile(int rt, int qt, int metric, const TileParams& p, dim3 grid, size_t smem, cudaStream_t s) {
    const bool cos = metric == ANNB_COSINE;
    if (rt == ANNB_F32 && qt == QT_F32) return cos ? launch_tile<0, QT_F32, MET_COS, EPI_SELECT>(p, grid, smem, s) : launch_tile<0, QT_F32, MET_L2, EPI_SELECT>(p, grid, smem, s);
    if (rt == ANNB_BF16 && qt == QT_F32) return cos ? launch_tile<1, QT_F32, MET_COS, EPI_SELECT>(p, grid, smem, s) : launch_tile<1, QT_F32, MET_L2, EPI_SELECT>(p, grid, smem, s);
    if (rt == ANNB_BF16 && qt == QT_BF16) return cos ? launch_tile<1, QT_BF16, MET_COS, EPI_SELECT>(p, grid, smem, s) : launch_tile<1, QT_BF16, MET_L2, EPI_SELECT>(p, grid, smem, s);
    if (rt == ANNB_SQ8 && qt == QT_I8) return cos ? launch_tile<2, QT_I8, MET_COS, EPI_SELECT>(p, grid, smem, s) : launch_tile<2, QT_I8, MET_L2, EPI_SELECT>(p, grid, smem, s);
    return fail(ANNB_ERR_INVALID_ARGUMENT, "unsupported (row, query) type pair");
}

template <int RT, int QT, int MET>
static int launch_scan_t(const ScanParams& p, uint32_t grid, size_t smem, cudaStream_t s) {
    auto kern = ivf_scan_kernel<RT, QT, MET>;
    if (smem > 227 * 1024) return fail(ANNB_ERR_UNSUPPORTED, "ivf_scan_kernel: dim/k too large for shared memory (DimTooHighForSharedMemory)");
    ANNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, SCAN_WARPS * 32, smem, s>>>(p);
    ANNB_CUDA_CHECK(cudaGetLastError());
    return ANNB_OK;
}
static int launch_scan(int rt, int qt, int metric, const ScanParams& p, uint32_t grid, size_t smem, cudaStream_t s) {
    const bool cos = metric == ANNB_COSINE;
    if (rt == ANNB_F32 && qt == QT_F32) return cos ? launch_scan_t<0, QT_F32, MET_COS>(p, grid, smem, s) : launch_scan_t<0, QT_F32, MET_L2>(p, grid, smem, s);
    if (rt == ANNB_BF16 && qt == QT_F32) return cos ? launch_scan_t<1, QT_F32, MET_COS>(p, grid, smem, s) : launch_scan_t<1, QT_F32, MET_L2>(p, grid, smem, s);
    if (rt == ANNB_BF16 && qt == QT_BF16) return cos ? launch_scan_t<1, QT_BF16, MET_COS>(p, grid, smem, s) : launch_scan_t<1, QT_BF16, MET_L2>(p, grid, smem, s);
    if (rt == ANNB_SQ8 && qt == QT_I8) return cos ? launch_scan_t<2, QT_I8, MET_COS>(p, grid, smem, s) : launch_scan_t<2, QT_I8, MET_L2>(p, grid, smem, s);
    return fail(ANNB_ERR_INVALID_ARGUMENT, "unsupported (row, query) type pair");
}

template <int RT, int QT, int MET>
static int launch_list_scan_t(const ListScanParams& p, uint32_t grid, size_t smem, cudaStream_t s) {
    auto kern = ivf_list_kernel<RT, QT, MET>;
    if (smem > 227 * 1024) return fail(ANNB_ERR_UNSUPPORTED, "ivf_list_kernel: dim/k too large for shared memory (DimTooHighForSharedMemory)");
    ANNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, TILE_THREADS, smem, s>>>(p);
    ANNB_CUDA_CHECK(cudaGetLastError());
    return ANNB_OK;
}
static int launch_list_scan(int rt, int qt, int metric, const ListScanParams& p, uint32_t grid, size_t smem, cudaStream_t s) {
    const bool cos = metric == ANNB_COSINE;
    if (rt == ANNB_F32 && qt == QT_F32) return cos ? launch_list_scan_t<0, QT_F32, MET_COS>(p, grid, smem, s) : launch_list_scan_t<0, QT_F32, MET_L2>(p, grid, smem, s);
    if (rt == ANNB_BF16 && qt == QT_F32) return cos ? launch_list_scan_t<1, QT_F32, MET_COS>(p, grid, smem, s) : launch_list_scan_t<1, QT_F32, MET_L2>(p, grid, smem, s);
    if (rt == ANNB_BF16 && qt == QT_BF16) return cos ? launch_list_scan_t<1, QT_BF16, MET_COS>(p, grid, smem, s) : launch_list_scan_t<1, QT_BF16, MET_L2>(p, grid, smem, s);
    if (rt == ANNB_SQ8 && qt == QT_I8) return cos ? launch_list_scan_t<2, QT_I8, MET_COS>(p, grid, smem, s) : launch_list_scan_t<2, QT_I8, MET_L2>(p, grid, smem, s);
    return fail(ANNB_ERR_INVALID_ARGUMENT, "unsupported (row, query) type pair");
}

static int run_finalize(annb_index* ix, const uint64_t* keys, uint32_t parts, uint32_t kc, uint32_t k_out, uint64_t nq,
                        const uint64_t* id_map, uint64_t id_base, const uint64_t* row_map, uint64_t* d_ids, float* d_dist,
                        uint32_t* d_cnt, cudaStream_t s, const uint32_t* parts_used = nullptr) {
    FinalizeParams f{};
    f.parts_used = parts_used;
    f.part_keys = keys; f.parts = parts; f.kc = kc; f.k = k_out;
    f.nsort = next_pow2(std::max(parts * kc, k_out));
    f.nq = nq; f.id_map = id_map; f.id_base = id_base; f.row_map = row_map;
    f.out_ids = d_ids; f.out_dist = d_dist; f.out_counts = d_cnt;
    size_t smem = static_cast<size_t>(f.nsort) * 8;
    if (smem > 200 * 1024) return fail(ANNB_ERR_UNSUPPORTED, "finalize: parts * k too large");
    ANNB_CUDA_CHECK(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    finalize_kernel<<<static_cast<uint32_t>(nq), f.nsort >= 2048 ? 512 : 128, smem, s>>>(f);
    ANNB_CUDA_CHECK(cudaGetLastError());
    ix->stat_launches++;
    return ANNB_OK;
}

// ---------------------------------------------------------------------------
// query preparation
// ---------------------------------------------------------------------------
struct PreparedQueries {
    const uint8_t* scan = nullptr;  // rows the distance kernels consume
    uint32_t scan_bytes = 0;
    int qt = QT_F32;
    int bf16_self = 0;
    const float* route = nullptr;  // f32 rows used for centroid ranking (IVF)
    uint32_t route_ld = 0;         // floats
};

// External f32 queries [nq][dim] (device) -> padded / encoded forms.
static int prepare_external(annb_index* ix, const float* d_q, uint64_t nq, PreparedQueries* out, cudaStream_t s) {
    const uint32_t dim = ix->dim;
    const uint32_t ld = round_up(dim * 4u, 16u) / 4u;
    const float* f32q = d_q;
    if (ld != dim || ix->dtype == ANNB_SQ8) {  // SQ8 may normalise in place -> always work on a copy
        ANNB_TRY(ix->s_qpad.ensure(nq * ld * 4ull));
        pad_rows_kernel<<<grid_for(nq * ld * 4ull, 256), 256, 0, s>>>(reinterpret_cast<const uint8_t*>(d_q), dim * 4, ix->s_qpad.as<uint8_t>(), ld * 4, nq);
        ANNB_CUDA_CHECK(cudaGetLastError());
        ix->stat_launches++;
        f32q = ix->s_qpad.as<float>();
    }
    out->route = f32q;
    out->route_ld = ld;
    out->bf16_self = 0;
    if (ix->dtype == ANNB_SQ8) {
        // ExhaustiveSq8Index::query (exhaustive_sq8.rs:181-192): normalise (cosine), encode with the codebook
        float* w = ix->s_qpad.as<float>();
        if (ix->metric == ANNB_COSINE) {
            normalise_rows_f32_kernel<<<grid_for(nq, 128, 1u << 30), 128, 0, s>>>(w, ld, dim, nq);
            ANNB_CUDA_CHECK(cudaGetLastError());
            ix->stat_launches++;
        }
        const uint32_t cb = round_up(dim, 16u);
        ANNB_TRY(ix->s_qcodes.ensure(nq * static_cast<uint64_t>(cb)));
        sq8_encode_kernel<<<grid_for(nq * cb, 256), 256, 0, s>>>(w, ld, dim, ix->d_scales, ix->s_qcodes.as<int8_t>(), cb, nq);
        ANNB_CUDA_CHECK(cudaGetLastError());
        ix->stat_launches++;
        out->scan = ix->s_qcodes.as<uint8_t>();
        out->scan_bytes = cb;
        out->qt = QT_I8;
    } else {
        out->scan = reinterpret_cast<const uint8_t*>(f32q);
        out->scan_bytes = ld * 4;
        out->qt = QT_F32;
    }
    return ANNB_OK;
}

// Self queries: stored rows [pos_begin, pos_end) act as queries in their own dtype.
static int prepare_self(annb_index* ix, uint64_t pos_begin, uint64_t nq, bool need_route, PreparedQueries* out, cudaStream_t s) {
    out->scan = ix->d_rows + pos_begin * ix->row_bytes;
    out->scan_bytes = ix->row_bytes;
    out->qt = ix->dtype == ANNB_F32 ? QT_F32 : (ix->dtype == ANNB_BF16 ? QT_BF16 : QT_I8);
    out->bf16_self = ix->dtype == ANNB_BF16;
    out->route = nullptr;
    if (!need_route) return ANNB_OK;
    const uint32_t ld = round_up(ix->dim * 4u, 16u) / 4u;
    out->route_ld = ld;
    if (ix->dtype == ANNB_F32) {
        out->route = reinterpret_cast<const float*>(out->scan);
        out->route_ld = ix->row_bytes / 4;
        return ANNB_OK;
    }
    ANNB_TRY(ix->s_route.ensure(nq * ld * 4ull));
    float* w = ix->s_route.as<float>();
    if (ix->dtype == ANNB_BF16)  // query_bf16: decode for routing (ivf_bf16.rs:456)
        decode_bf16_kernel<<<grid_for(nq * ld, 256), 256, 0, s>>>(reinterpret_cast<const uint16_t*>(out->scan), ix->row_bytes / 2, w, ld, ix->dim, nq);
    else  // query_quantised: decode codes for routing (ivf_sq8.rs:409)
        sq8_decode_kernel<<<grid_for(nq * ld, 256), 256, 0, s>>>(reinterpret_cast<const int8_t*>(out->scan), ix->row_bytes, ix->d_scales, w, ld, ix->dim, nq);
    ANNB_CUDA_CHECK(cudaGetLastError());
    ix->stat_launches++;
    out->route = w;
    return ANNB_OK;
}

// ---------------------------------------------------------------------------
// Status words.  A search batch is enqueued optimistically (tensor-core ranking / pre-selection); what may force a
// second look -- a probe set that outgrew its ranked prefix, queries that failed the coverage certificate -- is written
// by the kernels into two small device blocks and read back ONCE per batch into a pinned host block, together with the
// batch's results when the caller wants them on the host (no extra synchronisation on that path).
// ---------------------------------------------------------------------------
struct CoreState {
    bool tensor = false;    // the batch went through a tensor-path kernel with a coverage certificate
    bool routed = false;    // centroid ranking ran in this call (probe statistics and the overflow flag are meaningful)
    int stage = -1;         // IVF ranking stage used: 3 tensor cores, 2 fused CUDA-core select, 1 dense + ranked prefix, 0 dense + full sort
    uint32_t pitch = 0;     // IVF: probe pitch of the lists in s_probes
};
struct HostStatus {          // layout of ix->h_status (pinned)
    uint32_t overflow, export_overflow;
    unsigned long long scanned, probed, local;
    uint32_t n_unc, pad1;
    unsigned long long tiles;      // tensor-core IVF scan: 128-row tiles executed by the batch's tasks
};

static int enqueue_status_read(annb_index* ix, const CoreState& cs, cudaStream_t s) {
    HostStatus* h = reinterpret_cast<HostStatus*>(ix->h_status);
    std::memset(h, 0, sizeof(HostStatus));
    if (cs.routed) ANNB_CUDA_CHECK(cudaMemcpyAsync(h, ix->s_flags.p, 32, cudaMemcpyDeviceToHost, s));
    if (cs.routed && cs.tensor) ANNB_CUDA_CHECK(cudaMemcpyAsync(&h->tiles, ix->s_flags.as<uint8_t>() + 32, 8, cudaMemcpyDeviceToHost, s));
    if (cs.tensor && ix->s_uncert.p) ANNB_CUDA_CHECK(cudaMemcpyAsync(&h->n_unc, ix->s_uncert.p, 4, cudaMemcpyDeviceToHost, s));
    return ANNB_OK;
}
static bool wants_status(const annb_index* ix, const CoreState& cs) {
    return cs.routed || (cs.tensor && ix->opt_cert_fallback && ix->opt_cert_eps != 0.f && ix->shard_bound == nullptr);
}

// Calls on one handle share its scratch buffers; the mutex orders the host side, this orders the device side when
// consecutive calls use different streams.
static int order_after_previous(annb_index* ix, cudaStream_t s) {
    if (ix->last_event_valid && s != ix->last_stream) ANNB_CUDA_CHECK(cudaStreamWaitEvent(s, ix->last_event, 0));
    return ANNB_OK;
}
static int mark_call_done(annb_index* ix, cudaStream_t s) {
    if (!ix->last_event) ANNB_CUDA_CHECK(cudaEventCreateWithFlags(&ix->last_event, cudaEventDisableTiming));
    ANNB_CUDA_CHECK(cudaEventRecord(ix->last_event, s));
    ix->last_event_valid = true;
    ix->last_stream = s;
    return ANNB_OK;
}

// ---------------------------------------------------------------------------
// flat search core (device pointers, asynchronous on s)
// ---------------------------------------------------------------------------
static int flat_simt(annb_index* ix, const PreparedQueries& pq, uint64_t nq, uint32_t k, uint32_t kk, uint64_t* d_ids, float* d_dist,
                     uint32_t* d_cnt, cudaStream_t s, bool is_fallback = false);

static int flat_enqueue(annb_index* ix, const PreparedQueries& pq, uint64_t nq, uint32_t k, uint64_t* d_ids, float* d_dist,
                        uint32_t* d_cnt, cudaStream_t s, CoreState* cs) {
    const uint32_t kk = static_cast<uint32_t>(std::min<uint64_t>(k, ix->n));
    if (kk == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "k must be >= 1");
    bool use_tc = false;
    if (ix->opt_path != ANNB_PATH_SIMT) {
        use_tc = tc_flat_supported(ix, pq.qt, kk);
        if (!use_tc && ix->opt_path == ANNB_PATH_TENSOR)
            return fail(ANNB_ERR_UNSUPPORTED, "tensor path requested but this (dtype, dim, k) is not covered by it yet");
    }
    if (use_tc) {
        ix->stat_last_path = ANNB_PATH_TENSOR;
        cs->tensor = true;
        return tc_flat_search(ix, pq.scan, pq.scan_bytes, pq.qt, pq.bf16_self, nq, kk, k, d_ids, d_dist, d_cnt, s);
    }
    ix->stat_last_path = ANNB_PATH_SIMT;
    return flat_simt(ix, pq, nq, k, kk, d_ids, d_dist, d_cnt, s);
}

// Queries whose pre-selection could not be certified are recomputed on the exact path (rare) and scattered over their rows.
static int flat_fallback(annb_index* ix, const PreparedQueries& pq, uint32_t k, uint32_t n_unc, uint64_t* d_ids, float* d_dist, uint32_t* d_cnt,
                         cudaStream_t s) {
    const uint32_t kk = static_cast<uint32_t>(std::min<uint64_t>(k, ix->n));
    const uint32_t* list = ix->s_uncert.as<uint32_t>() + 1;
    PreparedQueries sub = pq;
    ANNB_TRY(ix->s_fbq.ensure(static_cast<uint64_t>(n_unc) * pq.scan_bytes));
    gather_rows_kernel<<<grid_for(static_cast<uint64_t>(n_unc) * (pq.scan_bytes >> 4), 256), 256, 0, s>>>(pq.scan, pq.scan_bytes, list, n_unc, ix->s_fbq.as<uint8_t>());
    ANNB_CUDA_CHECK(cudaGetLastError());
    sub.scan = ix->s_fbq.as<uint8_t>();
    ANNB_TRY(ix->s_fbi.ensure(static_cast<uint64_t>(n_unc) * k * 8));
    ANNB_TRY(ix->s_fbd.ensure(static_cast<uint64_t>(n_unc) * k * 4));
    ANNB_TRY(ix->s_fbc.ensure(static_cast<uint64_t>(n_unc) * 4));
    ANNB_TRY(flat_simt(ix, sub, n_unc, k, kk, ix->s_fbi.as<uint64_t>(), ix->s_fbd.as<float>(), ix->s_fbc.as<uint32_t>(), s, true));
    scatter_results_kernel<<<grid_for(static_cast<uint64_t>(n_unc) * k, 256), 256, 0, s>>>(list, n_unc, k, ix->s_fbi.as<uint64_t>(), ix->s_fbd.as<float>(),
                                                                                         ix->s_fbc.as<uint32_t>(), d_ids, d_dist, d_cnt);
    ANNB_CUDA_CHECK(cudaGetLastError());
    ix->stat_launches += 2;
    ix->stat_fallback_queries += n_unc;
    return ANNB_OK;
}

// Exact CUDA-core flat search (tile_kernel + finalize).
static int flat_simt(annb_index* ix, const PreparedQueries& pq, uint64_t nq, uint32_t k, uint32_t kk, uint64_t* d_ids, float* d_dist,
                     uint32_t* d_cnt, cudaStream_t s, bool is_fallback) {
    if (kk > 1024) return fail(ANNB_ERR_UNSUPPORTED, "k > 1024 is not supported");
    const uint32_t nsort = WarpSelect::sort_size(kk);
    const uint64_t q_tiles = ceil_div<uint64_t>(nq, CTA_QUERIES);
    uint32_t splits = ix->opt_db_splits > 0 ? static_cast<uint32_t>(ix->opt_db_splits)
                                            : static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(592, ceil_div<uint64_t>(4 * 148, q_tiles))));   // tiny batches (the tensor path's fallback): up to four waves of splits
    splits = static_cast<uint32_t>(std::min<uint64_t>(splits, std::max<uint64_t>(1, ix->n / 256)));
    splits = std::min<uint32_t>(splits, std::max<uint32_t>(1, 16384u / kk));   // finalize sorts splits * k keys in shared memory
    uint64_t rps = round_up<uint64_t>(ceil_div<uint64_t>(ix->n, splits), TILE_ROWS);
    splits = static_cast<uint32_t>(ceil_div<uint64_t>(ix->n, rps));
    ANNB_TRY(ix->s_keys.ensure(nq * splits * static_cast<uint64_t>(kk) * 8));
    TileParams p{};
    p.rows = ix->d_rows; p.n_rows = ix->n; p.row_bytes = ix->row_bytes;
    p.row_norms = ix->d_norms; p.row_norms_i = ix->d_norms_i;
    p.queries = pq.scan; p.q_bytes = pq.scan_bytes; p.nq = nq; p.dim = ix->dim; p.bf16_self = pq.bf16_self;
    p.k = kk; p.nsort = nsort; p.n_splits = splits; p.rows_per_split = rps; p.part_keys = ix->s_keys.as<uint64_t>();
    size_t smem = tile_kernel_smem(ix->row_bytes, pq.scan_bytes, nsort, true);
    {
        KernelTimer kt(ix, s, !is_fallback);
        ANNB_TRY(launch_select_tile(ix->dtype, pq.qt, ix->metric, p, dim3(static_cast<uint32_t>(q_tiles), splits), smem, s));
    }
    ix->stat_launches++;
    return run_finalize(ix, ix->s_keys.as<uint64_t>(), splits, kk, k, nq, nullptr, ix->id_base, nullptr, d_ids, d_dist, d_cnt, s);
}

// ---------------------------------------------------------------------------
// IVF search core
// ---------------------------------------------------------------------------
struct RouteOut { uint32_t* probes; uint32_t* n_probes; uint32_t pitch; };

static void fill_probe_params(annb_index* ix, ProbeParams& pp, uint64_t nq, uint32_t np, uint32_t kk, uint32_t pitch) {
    pp.nq = nq; pp.nlist = ix->nlist; pp.offsets = ix->d_offsets; pp.nprobe = np; pp.k = kk;
    pp.probes = ix->s_probes.as<uint32_t>(); pp.probe_pitch = pitch; pp.n_probes = ix->s_nprobes.as<uint32_t>();
    pp.overflow = ix->s_flags.as<uint32_t>();
    pp.stat_scanned = reinterpret_cast<unsigned long long*>(ix->s_flags.as<uint8_t>() + 8);
    pp.stat_probed = reinterpret_cast<unsigned long long*>(ix->s_flags.as<uint8_t>() + 16);
    pp.list_begin = ix->list_begin; pp.list_end = ix->list_end;
}

// Centroid ranking + probe expansion (src/cpu/ivf.rs:349-365) into s_probes / s_nprobes.  Stages, most optimistic first:
//   3  dense approximate values on the tensor cores, per query radix select of the nearest np + 32 cells, exact distances
//      for those, certified (distance, cell) prefix, probe walk;
//   2  exact fused select of a short ranked prefix on the CUDA cores (tile_kernel<EPI_SELECT>), probe walk;
//   1  exact dense matrix (get_centroids_dist / _prenorm arithmetic) + full per-query sort, probes cut at the prefix pitch;
//   0  the same with pitch = nlist (cannot overflow).
// Stages 3..1 raise the overflow flag (s_flags[0]) when some query needs more cells than the stage can rank (or certify);
// nothing is read back here -- the caller repeats the batch from the next stage if the flag turns out set.
static int ivf_route(annb_index* ix, const PreparedQueries& pq, uint64_t nq, uint32_t np, uint32_t kk, int stage_start, bool force_simt,
                     CoreState* cs, cudaStream_t s) {
    ANNB_TRY(ix->s_flags.ensure(64));
    ANNB_TRY(ix->s_nprobes.ensure(nq * 4));
    ANNB_CUDA_CHECK(cudaMemsetAsync(ix->s_flags.p, 0, 64, s));
    cs->routed = true;
    // ranked prefix length: room for probe expansion (>= nprobe + 16 cells) and pitch + 64 a power of two, so the fused
    // select's sort buffer is exactly full (smaller shared memory -> more resident CTAs)
    const uint32_t pitch_short = ix->nlist <= 1024 ? ix->nlist : std::min(ix->nlist, next_pow2(np + 16 + 64) - 64);
    if (stage_start >= 3 && !force_simt && ix->opt_path != ANNB_PATH_SIMT && ix->opt_ivf_tc_coarse && tc_coarse_supported(ix)) {
        const uint32_t p_tc = round_up(np + 32u, 32u);
        if (p_tc < ix->nlist && p_tc <= 480) {
            ANNB_TRY(ix->s_cdist.ensure(nq * static_cast<uint64_t>(p_tc) * 8));
            ANNB_TRY(ix->s_probes.ensure(nq * static_cast<uint64_t>(p_tc) * 4));
            ProbeParams pp{};
            fill_probe_params(ix, pp, nq, np, kk, p_tc);
            if (ix->opt_ivf_coarse_walk) {      // the select's warp applies the probe-expansion rule to its own certified prefix
                CoarseWalk w{pp.offsets, pp.nprobe, pp.k, pp.probes, pp.n_probes, pp.overflow, pp.stat_scanned, pp.stat_probed, pp.list_begin, pp.list_end};
                ANNB_TRY(tc_coarse_rank(ix, pq.route, pq.route_ld, nq, p_tc, ix->s_cdist.as<uint64_t>(), s, &w));
            } else {
                ANNB_TRY(tc_coarse_rank(ix, pq.route, pq.route_ld, nq, p_tc, ix->s_cdist.as<uint64_t>(), s));
                probe_walk_kernel<<<static_cast<uint32_t>(ceil_div<uint64_t>(nq, 128)), 128, 0, s>>>(ix->s_cdist.as<uint64_t>(), pp);
                ANNB_CUDA_CHECK(cudaGetLastError());
                ix->stat_launches++;
            }
            cs->stage = 3; cs->pitch = p_tc; ix->stat_coarse_path = 2;
            return ANNB_OK;
        }
    }
    // (wide prefixes make the fused select's sort buffers large and its pass rate high: above 64 ranks the dense matrix + sort wins)
    if (stage_start >= 2 && pitch_short < ix->nlist && (ix->opt_ivf_fast_probe == 1 ? pitch_short <= 64 : ix->opt_ivf_fast_probe != 0)) {
        const uint32_t pitch = pitch_short;
        const uint32_t nsort = WarpSelect::sort_size(pitch);
        const uint32_t cb = ix->cent_ld * 4, qb = pq.route_ld * 4;
        const size_t smem = tile_kernel_smem(cb, qb, nsort, true);
        if (smem <= 200 * 1024) {
            ANNB_TRY(ix->s_cdist.ensure(nq * static_cast<uint64_t>(pitch) * 8));
            ANNB_TRY(ix->s_probes.ensure(nq * static_cast<uint64_t>(pitch) * 4));
            TileParams p{};
            p.rows = reinterpret_cast<const uint8_t*>(ix->d_centroids); p.n_rows = ix->nlist; p.row_bytes = cb; p.row_norms = ix->d_centroid_norms;
            p.queries = reinterpret_cast<const uint8_t*>(pq.route); p.q_bytes = qb; p.nq = nq; p.dim = ix->dim;
            p.k = pitch; p.nsort = nsort; p.n_splits = 1; p.rows_per_split = round_up<uint64_t>(ix->nlist, TILE_ROWS); p.part_keys = ix->s_cdist.as<uint64_t>();
            dim3 grid(static_cast<uint32_t>(ceil_div<uint64_t>(nq, CTA_QUERIES)), 1);
            if (ix->metric == ANNB_L2) ANNB_TRY((launch_tile<0, QT_F32, MET_L2, EPI_SELECT>(p, grid, smem, s)));
            else if (ix->dtype == ANNB_SQ8) ANNB_TRY((launch_tile<0, QT_F32, MET_COS_PRENORM, EPI_SELECT>(p, grid, smem, s)));
            else ANNB_TRY((launch_tile<0, QT_F32, MET_COS, EPI_SELECT>(p, grid, smem, s)));
            ProbeParams pp{};
            fill_probe_params(ix, pp, nq, np, kk, pitch);
            probe_walk_kernel<<<static_cast<uint32_t>(ceil_div<uint64_t>(nq, 128)), 128, 0, s>>>(ix->s_cdist.as<uint64_t>(), pp);
            ANNB_CUDA_CHECK(cudaGetLastError());
            ix->stat_launches += 2;
            cs->stage = 2; cs->pitch = pitch; ix->stat_coarse_path = 1;
            return ANNB_OK;
        }
    }
    // query -> all centroids (dense), reference arithmetic of get_centroids_dist / _prenorm; full per-query sort
    const uint32_t nl2 = next_pow2(ix->nlist);
    // (the full sort of a centroid row lives in shared memory: tables of more than 16 384 cells are served by the ranked-prefix
    // stages above only -- a batch that needs more cells than the prefix holds, e.g. many tiny lists, is refused here)
    if (static_cast<size_t>(nl2) * 8 > 200 * 1024)
        return fail(ANNB_ERR_UNSUPPORTED, "nlist > 16384: this batch needs the full centroid ranking (a probe set outgrew the ranked prefix), which is not supported for tables this large");
    const uint32_t pitch = stage_start >= 1 ? pitch_short : ix->nlist;
    ANNB_TRY(ix->s_cdist.ensure(nq * static_cast<uint64_t>(ix->nlist) * 4));
    {
        TileParams p{};
        p.rows = reinterpret_cast<const uint8_t*>(ix->d_centroids); p.n_rows = ix->nlist; p.row_bytes = ix->cent_ld * 4;
        p.row_norms = ix->d_centroid_norms;
        p.queries = reinterpret_cast<const uint8_t*>(pq.route); p.q_bytes = pq.route_ld * 4; p.nq = nq; p.dim = ix->dim;
        p.dense_out = ix->s_cdist.as<float>();
        size_t smem = tile_kernel_smem(p.row_bytes, p.q_bytes, 0, false);
        dim3 grid(static_cast<uint32_t>(ceil_div<uint64_t>(nq, CTA_QUERIES)), 1);
        if (ix->metric == ANNB_L2) ANNB_TRY((launch_tile<0, QT_F32, MET_L2, EPI_DENSE>(p, grid, smem, s)));
        else if (ix->dtype == ANNB_SQ8) ANNB_TRY((launch_tile<0, QT_F32, MET_COS_PRENORM, EPI_DENSE>(p, grid, smem, s)));
        else ANNB_TRY((launch_tile<0, QT_F32, MET_COS, EPI_DENSE>(p, grid, smem, s)));
        ix->stat_launches++;
    }
    ANNB_TRY(ix->s_probes.ensure(nq * static_cast<uint64_t>(pitch) * 4));
    ProbeParams pp{};
    fill_probe_params(ix, pp, nq, np, kk, pitch);
    pp.cdist = ix->s_cdist.as<float>(); pp.nlist_pow2 = nl2;
    const size_t smem = static_cast<size_t>(nl2) * 8;
    ANNB_CUDA_CHECK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    probe_kernel<<<static_cast<uint32_t>(nq), 256, smem, s>>>(pp);
    ANNB_CUDA_CHECK(cudaGetLastError());
    ix->stat_launches++;
    cs->stage = pitch == ix->nlist ? 0 : 1; cs->pitch = pitch; ix->stat_coarse_path = 0;
    return ANNB_OK;
}

static int ivf_enqueue(annb_index* ix, const PreparedQueries& pq, uint64_t nq, uint32_t k, uint32_t nprobe, const uint64_t* row_map,
                       uint64_t* d_ids, float* d_dist, uint32_t* d_cnt, cudaStream_t s, CoreState* cs, int stage_start = 3, bool force_simt = false,
                       const uint32_t* preset_probes = nullptr, const uint32_t* preset_nprobes = nullptr, uint32_t preset_pitch = 0,
                       const RouteOut* route_out = nullptr) {
    // preset_*: probe lists already computed for these queries (the exact fallback of the tensor path re-uses the
    // parent call's ranking instead of ranking the centroids again; sharded searches route a slice of the batch per rank).
    // route_out: stop after the routing stage and export the probe lists.
    const uint32_t kk = static_cast<uint32_t>(std::min<uint64_t>(k, ix->n_total));
    if (kk == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "k must be >= 1");
    if (kk > 1024) return fail(ANNB_ERR_UNSUPPORTED, "k > 1024 is not supported");
    // nprobe default and clamp: src/cpu/ivf.rs:345-347
    uint32_t np = nprobe ? nprobe : std::max<uint32_t>(1, static_cast<uint32_t>(std::sqrt(static_cast<double>(ix->nlist))));
    np = std::min(np, ix->nlist);

    uint32_t pitch = preset_pitch;
    if (!preset_probes) {
        ANNB_TRY(ivf_route(ix, pq, nq, np, kk, stage_start, force_simt, cs, s));
        pitch = cs->pitch;
    }
    const uint32_t* d_probes = preset_probes ? preset_probes : ix->s_probes.as<uint32_t>();
    const uint32_t* d_nprobes = preset_probes ? preset_nprobes : ix->s_nprobes.as<uint32_t>();
    if (route_out != nullptr) {
        // the export shares the overflow flag with the ranking: a probe set that does not fit the caller's pitch can only be
        // reported, a ranking that ran out of prefix is repeated from the next stage (see ivf_finish_route)
        export_probes_kernel<<<grid_for(nq * route_out->pitch, 256, 1u << 30), 256, 0, s>>>(d_probes, pitch, d_nprobes, nq, route_out->probes, route_out->pitch,
                                                                                      route_out->n_probes, ix->s_flags.as<uint32_t>() + 1);
        ANNB_CUDA_CHECK(cudaGetLastError());
        ix->stat_launches++;
        return ANNB_OK;
    }
    // list scan.  A batch that probes every list many times goes list-major (one staged list tile serves many queries);
    // small batches keep the query-major streaming kernel (one warp per query part).
    const uint64_t n_local_lists = std::max<uint32_t>(1, ix->list_end - ix->list_begin);
    // the merge kernels sort all per-(query, rank) lists of a query in shared memory: very wide probe sets go query-major
    // (both kernels sort a power-of-two padded array, so the guard is on the padded size)
    const auto sort_fits = [](uint64_t entries) { return entries <= (1u << 30) && static_cast<uint64_t>(next_pow2(static_cast<uint32_t>(std::max<uint64_t>(entries, 64)))) * 8 <= 200 * 1024; };
    const bool merge_fits_simt = sort_fits(static_cast<uint64_t>(pitch) * kk);
    const bool merge_fits_tc = sort_fits(static_cast<uint64_t>(pitch) * 2 * tc_ivf_kprime(ix, kk));
    const bool list_major = merge_fits_simt &&
                            (ix->opt_ivf_list_major == 1 || (ix->opt_ivf_list_major < 0 && nq * static_cast<uint64_t>(np) * 8 >= n_local_lists));
    if (list_major) {
        const bool use_tc = !force_simt && ix->opt_path != ANNB_PATH_SIMT && merge_fits_tc && tc_ivf_supported(ix, pq.qt, kk);
        if (!use_tc && !force_simt && ix->opt_path == ANNB_PATH_TENSOR)
            return fail(ANNB_ERR_UNSUPPORTED, "tensor path requested but this IVF (dtype, dim, k, query type) is not covered by it yet");
        ix->stat_last_path = use_tc ? ANNB_PATH_TENSOR : ANNB_PATH_SIMT;
        const uint32_t nsort = WarpSelect::sort_size(kk);
        const uint64_t slots = nq * static_cast<uint64_t>(pitch);
        if (!use_tc) {
            ANNB_TRY(ix->s_keys.ensure(slots * kk * 8));
            ANNB_CUDA_CHECK(cudaMemsetAsync(ix->s_keys.p, 0xFF, slots * kk * 8, s));   // unvisited (query, rank) slots stay sentinels
        }
        // scratch: cnt | cursor | pair_off | task_off | task_counter | pairs
        const size_t nl = ix->nlist;
        const size_t hdr = (4 * (nl + 1) + 4) * sizeof(uint32_t);
        const uint64_t max_tasks_tc = ceil_div<uint64_t>(slots, 128) + n_local_lists;
        const size_t pairs_bytes = round_up<size_t>(slots * sizeof(uint2), 16);
        ANNB_TRY(ix->s_pairs.ensure(hdr + pairs_bytes + (use_tc ? max_tasks_tc * 32 : 0)));
        uint32_t* w = ix->s_pairs.as<uint32_t>();
        ANNB_CUDA_CHECK(cudaMemsetAsync(w, 0, hdr, s));
        PairParams pp{};
        pp.probes = d_probes; pp.probe_pitch = pitch; pp.n_probes = d_nprobes; pp.nq = nq;
        pp.nlist = ix->nlist; pp.list_begin = ix->list_begin; pp.list_end = ix->list_end;
        pp.cnt = w; pp.cursor = w + (nl + 1); pp.pair_off = w + 2 * (nl + 1); pp.task_off = w + 3 * (nl + 1); pp.task_counter = w + 4 * (nl + 1);
        pp.pairs = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(w) + hdr);
        pp.group = use_tc ? 128u : static_cast<uint32_t>(CTA_QUERIES);
        pp.tasks = use_tc ? reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(w) + hdr + pairs_bytes) : nullptr;
        pp.offsets = ix->d_offsets; pp.shard_row0 = ix->shard_row0;
        pp.order = (use_tc && ix->opt_ivf_task_order) ? ix->d_list_order : nullptr;
        pp.stat_tiles = (use_tc && !preset_probes && ix->s_flags.p) ? reinterpret_cast<unsigned long long*>(ix->s_flags.as<uint8_t>() + 32) : nullptr;
        const uint32_t g = static_cast<uint32_t>(ceil_div<uint64_t>(slots, 256));
        ivf_count_pairs_kernel<<<g, 256, 0, s>>>(pp);
        ivf_pair_offsets_kernel<<<1, 1024, 0, s>>>(pp);
        ivf_fill_pairs_kernel<<<g, 256, 0, s>>>(pp);
        ANNB_CUDA_CHECK(cudaGetLastError());
        ix->stat_launches += 3;
        if (use_tc) {
            cs->tensor = true;
            return tc_ivf_scan(ix, pq.scan, pq.scan_bytes, nq, kk, k, pitch, pp.pair_off, pp.task_off, pp.pairs, pp.task_counter, max_tasks_tc,
                               d_nprobes, row_map, d_ids, d_dist, d_cnt, s, pp.tasks);
        }
        ListScanParams lp{};
        lp.rows = ix->d_rows; lp.row_bytes = ix->row_bytes; lp.row_norms = ix->d_norms; lp.row_norms_i = ix->d_norms_i;
        lp.queries = pq.scan; lp.q_bytes = pq.scan_bytes; lp.dim = ix->dim; lp.bf16_self = pq.bf16_self;
        lp.offsets = ix->d_offsets; lp.shard_row0 = ix->shard_row0; lp.nlist = ix->nlist;
        lp.pair_off = pp.pair_off; lp.task_off = pp.task_off; lp.pairs = pp.pairs; lp.task_counter = pp.task_counter;
        lp.probe_pitch = pitch; lp.k = kk; lp.nsort = nsort; lp.part_keys = ix->s_keys.as<uint64_t>();
        const size_t smem = list_kernel_smem(ix->row_bytes, pq.scan_bytes, nsort);
        const uint64_t max_tasks = ceil_div<uint64_t>(slots, CTA_QUERIES) + n_local_lists;
        const uint32_t ctas_per_sm = static_cast<uint32_t>(std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / std::max<size_t>(smem, 1))));
        const uint32_t grid = static_cast<uint32_t>(std::min<uint64_t>(max_tasks, 148ull * ctas_per_sm));
        {
            KernelTimer kt(ix, s);
            ANNB_TRY(launch_list_scan(ix->dtype, pq.qt, ix->metric, lp, grid, smem, s));
        }
        ix->stat_launches++;
        return run_finalize(ix, ix->s_keys.as<uint64_t>(), pitch, kk, k, nq, ix->d_original_ids, 0, row_map, d_ids, d_dist, d_cnt, s, d_nprobes);
    }
    ix->stat_last_path = ANNB_PATH_SIMT;
    // warps wanted per query so that a small batch still fills the machine (148 SMs x 32 warps)
    const uint64_t want = std::max<uint64_t>(1, std::min<uint64_t>(1024, ceil_div<uint64_t>(148ull * 32, std::max<uint64_t>(nq, 1))));
    uint32_t parts = ix->opt_scan_parts > 0 ? static_cast<uint32_t>(ix->opt_scan_parts) : static_cast<uint32_t>(std::min<uint64_t>(want, 32));
    parts = std::max(1u, std::min(parts, np));
    // very small batches (the tensor path's exact fallback): also cut every list into row segments
    uint32_t subs = ix->opt_scan_parts > 0 ? 1u : static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(32, want / parts)));
    while (subs > 1 && static_cast<uint64_t>(parts) * subs * kk * 8 > 32 * 1024) subs--;   // finalize sorts parts * subs * k keys per query: keep that a 4096-key sort
    const uint32_t nsort = WarpSelect::sort_size(kk);
    ANNB_TRY(ix->s_keys.ensure(nq * parts * subs * static_cast<uint64_t>(kk) * 8));
    {
        ScanParams sp{};
        sp.rows = ix->d_rows; sp.row_bytes = ix->row_bytes; sp.row_norms = ix->d_norms; sp.row_norms_i = ix->d_norms_i;
        sp.queries = pq.scan; sp.q_bytes = pq.scan_bytes; sp.nq = nq; sp.dim = ix->dim; sp.bf16_self = pq.bf16_self;
        sp.probes = d_probes; sp.probe_pitch = pitch; sp.n_probes = d_nprobes;
        sp.offsets = ix->d_offsets; sp.list_begin = ix->list_begin; sp.list_end = ix->list_end; sp.shard_row0 = ix->shard_row0;
        sp.parts = parts; sp.subs = subs; sp.k = kk; sp.nsort = nsort; sp.part_keys = ix->s_keys.as<uint64_t>();
        size_t smem = scan_kernel_smem(ix->row_bytes, pq.scan_bytes, nsort);
        uint32_t grid = static_cast<uint32_t>(ceil_div<uint64_t>(nq * parts * subs, SCAN_WARPS));
        {
            KernelTimer kt(ix, s, !force_simt);   // (the exact fallback of a tensor-path call is not the dominant kernel)
            if (tc_stream_supported(ix)) {
                StreamScanArgs a{pq.scan, pq.scan_bytes, nq, pq.qt, pq.bf16_self, d_probes, pitch, d_nprobes, parts, subs, kk, nsort, ix->s_keys.as<uint64_t>()};
                ANNB_TRY(tc_stream_scan(ix, a, s));
            } else {
                ANNB_TRY(launch_scan(ix->dtype, pq.qt, ix->metric, sp, grid, smem, s));
            }
        }
        ix->stat_launches++;
    }
    // merge parts, map list-order positions to original ids (src/cpu/ivf.rs:383-389)
    return run_finalize(ix, ix->s_keys.as<uint64_t>(), parts * subs, kk, k, nq, ix->d_original_ids, 0, row_map, d_ids, d_dist, d_cnt, s);
}

// Exact fallback of a tensor-path IVF batch: the uncertified queries go through the CUDA-core pipeline again, re-using
// their probe lists ([n_unc][pitch] cell ids followed by [n_unc] probe counts).
static int ivf_fallback(annb_index* ix, const PreparedQueries& pq, uint32_t k, uint32_t nprobe, uint32_t n_unc, const uint32_t* d_probes,
                        const uint32_t* d_nprobes, uint32_t pitch, uint64_t* d_ids, float* d_dist, uint32_t* d_cnt, cudaStream_t s) {
    const uint32_t* list = ix->s_uncert.as<uint32_t>() + 1;
    PreparedQueries sub = pq;
    ANNB_TRY(ix->s_fbq.ensure(static_cast<uint64_t>(n_unc) * pq.scan_bytes));
    gather_rows_kernel<<<grid_for(static_cast<uint64_t>(n_unc) * (pq.scan_bytes >> 4), 256), 256, 0, s>>>(pq.scan, pq.scan_bytes, list, n_unc, ix->s_fbq.as<uint8_t>());
    sub.scan = ix->s_fbq.as<uint8_t>();
    ANNB_TRY(ix->s_fbr.ensure(static_cast<uint64_t>(n_unc) * (pitch + 1) * 4));
    uint32_t* fb_probes = ix->s_fbr.as<uint32_t>();
    uint32_t* fb_nprobes = fb_probes + static_cast<uint64_t>(n_unc) * pitch;
    gather_u32_rows_kernel<<<grid_for(static_cast<uint64_t>(n_unc) * pitch, 256), 256, 0, s>>>(d_probes, pitch, list, n_unc, fb_probes);
    gather_u32_rows_kernel<<<grid_for(n_unc, 256), 256, 0, s>>>(d_nprobes, 1, list, n_unc, fb_nprobes);
    ANNB_CUDA_CHECK(cudaGetLastError());
    sub.route = nullptr;
    ANNB_TRY(ix->s_fbi.ensure(static_cast<uint64_t>(n_unc) * k * 8));
    ANNB_TRY(ix->s_fbd.ensure(static_cast<uint64_t>(n_unc) * k * 4));
    ANNB_TRY(ix->s_fbc.ensure(static_cast<uint64_t>(n_unc) * 4));
    // the list of uncertified queries lives in s_uncert, which the nested call does not touch (it stays on the CUDA-core path)
    CoreState nested;
    ANNB_TRY(ivf_enqueue(ix, sub, n_unc, k, nprobe, nullptr, ix->s_fbi.as<uint64_t>(), ix->s_fbd.as<float>(), ix->s_fbc.as<uint32_t>(), s, &nested, 0, true,
                         fb_probes, fb_nprobes, pitch));
    scatter_results_kernel<<<grid_for(static_cast<uint64_t>(n_unc) * k, 256), 256, 0, s>>>(list, n_unc, k, ix->s_fbi.as<uint64_t>(), ix->s_fbd.as<float>(),
                                                                                         ix->s_fbc.as<uint32_t>(), d_ids, d_dist, d_cnt);
    ANNB_CUDA_CHECK(cudaGetLastError());
    ix->stat_launches += 4;
    ix->stat_fallback_queries += n_unc;
    ix->stat_last_path = ANNB_PATH_TENSOR;
    return ANNB_OK;
}

// ids[i] = map[ids[i]] (positions -> original ids); out-of-range entries (padding) are left alone
static __global__ void map_ids_kernel(uint64_t* __restrict__ ids, uint64_t count, const uint64_t* __restrict__ map, uint64_t n) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i < count && ids[i] < n) ids[i] = map[ids[i]];
}

constexpr uint64_t QUERY_BATCH = 16384;

// One batch, from prepared queries to final results in the device buffers d_*: enqueue, read the status words back once
// (together with whatever `copy_out` enqueues -- the host entry points copy the results in the same synchronisation), and
// only if they say so repeat the ranking from the next stage / recompute the uncertified queries and copy again.
// preset_*: probe lists supplied by the caller (sharded search after a probe exchange).
template <typename CopyOut>
static int run_batch(annb_index* ix, bool ivf, const PreparedQueries& pq, uint64_t nb, uint32_t k, uint32_t nprobe, uint64_t* d_ids, float* d_dist,
                     uint32_t* d_cnt, cudaStream_t s, bool may_sync, CopyOut copy_out, const uint32_t* preset_probes = nullptr,
                     const uint32_t* preset_nprobes = nullptr, uint32_t preset_pitch = 0) {
    int stage = 3;
    for (;;) {
        CoreState cs;
        if (ivf) ANNB_TRY(ivf_enqueue(ix, pq, nb, k, nprobe, nullptr, d_ids, d_dist, d_cnt, s, &cs, stage, false, preset_probes, preset_nprobes, preset_pitch));
        else ANNB_TRY(flat_enqueue(ix, pq, nb, k, d_ids, d_dist, d_cnt, s, &cs));
        ANNB_TRY(copy_out());
        if (!may_sync || !wants_status(ix, cs)) {
            if (may_sync) ANNB_CUDA_CHECK(cudaStreamSynchronize(s));
            return ANNB_OK;
        }
        ANNB_TRY(enqueue_status_read(ix, cs, s));
        ANNB_CUDA_CHECK(cudaStreamSynchronize(s));
        const HostStatus h = *reinterpret_cast<const HostStatus*>(ix->h_status);
        if (cs.routed && h.overflow && cs.stage > 0) { stage = cs.stage - 1; continue; }   // rare: some probe set outgrew the ranked / certified prefix
        if (cs.routed) {
            ix->stat_scanned += static_cast<int64_t>(h.scanned);
            ix->stat_probed += static_cast<int64_t>(h.probed);
            ix->stat_scanned_local += static_cast<int64_t>(h.local);
            ix->stat_tc_tiles += static_cast<int64_t>(h.tiles);
        }
        const uint32_t n_unc = (cs.tensor && ix->opt_cert_fallback && ix->opt_cert_eps != 0.f && ix->shard_bound == nullptr) ? h.n_unc : 0u;
        ix->stat_uncertified = cs.tensor ? static_cast<int64_t>(h.n_unc) : 0;
        if (!ivf && cs.tensor && n_unc >= 8 && static_cast<uint64_t>(n_unc) * 50 > nb) ix->tc_escalate = std::min(ix->tc_escalate + 1, 2);   // see tc_flat_search: k' = 32, then wide-k mode, from the next batch on
        if (n_unc == 0) return ANNB_OK;
        if (ivf) ANNB_TRY(ivf_fallback(ix, pq, k, nprobe, n_unc, preset_probes ? preset_probes : ix->s_probes.as<uint32_t>(),
                                       preset_probes ? preset_nprobes : ix->s_nprobes.as<uint32_t>(), preset_probes ? preset_pitch : cs.pitch, d_ids, d_dist, d_cnt, s));
        else ANNB_TRY(flat_fallback(ix, pq, k, n_unc, d_ids, d_dist, d_cnt, s));
        ANNB_TRY(copy_out());
        ANNB_CUDA_CHECK(cudaStreamSynchronize(s));
        return ANNB_OK;
    }
}

// Routing only (sharded searches): probe lists of one batch exported at the caller's pitch; repeats from the next ranking
// stage when the optimistic one ran out of (certified) prefix.
static int route_batch(annb_index* ix, const PreparedQueries& pq, uint64_t nb, uint32_t k, uint32_t nprobe, const RouteOut& ro, cudaStream_t s) {
    for (int stage = 3;;) {
        CoreState cs;
        ANNB_TRY(ivf_enqueue(ix, pq, nb, k, nprobe, nullptr, nullptr, nullptr, nullptr, s, &cs, stage, false, nullptr, nullptr, 0, &ro));
        if (ix->opt_async_dev) return ANNB_OK;
        ANNB_TRY(enqueue_status_read(ix, cs, s));
        ANNB_CUDA_CHECK(cudaStreamSynchronize(s));
        const HostStatus h = *reinterpret_cast<const HostStatus*>(ix->h_status);
        if (h.overflow && cs.stage > 0) { stage = cs.stage - 1; continue; }
        if (h.export_overflow) return fail(ANNB_ERR_UNSUPPORTED, "a query needs more probed lists than the caller's probe pitch");
        return ANNB_OK;
    }
}

// Host-buffer driver shared by the four host entry points.
//   mode 0: external queries (host or device f32 [nq][dim]);  mode 1: self queries [pos_begin, pos_begin + nq)
static int search_host(annb_index* ix, bool ivf, int mode, const float* queries, uint64_t pos_begin, uint64_t nq, uint32_t k,
                       uint32_t nprobe, int scatter, uint64_t* out_ids, float* out_dist, uint32_t* out_counts) {
    if (!ix) return fail(ANNB_ERR_INVALID_ARGUMENT, "null index");
    if (ivf != ix->is_ivf) return fail(ANNB_ERR_INVALID_ARGUMENT, ivf ? "not an IVF index" : "not a flat index");
    if (!out_ids || (mode == 0 && !queries && nq)) return fail(ANNB_ERR_INVALID_ARGUMENT, "null buffer");
    if (k == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "k must be >= 1");
    ANNB_DEVICE(ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    cudaStream_t s = ix->stream;
    ANNB_TRY(order_after_previous(ix, s));
    ix->stat_scanned = 0;
    ix->stat_probed = 0;
    ix->stat_scanned_local = 0;
    ix->stat_tc_tiles = 0;
    for (uint64_t b0 = 0; b0 < nq; b0 += QUERY_BATCH) {
        const uint64_t nb = std::min<uint64_t>(QUERY_BATCH, nq - b0);
        PreparedQueries pq;
        if (mode == 0) {
            ANNB_TRY(ix->s_tmp.ensure(nb * ix->dim * 4ull));
            ANNB_CUDA_CHECK(cudaMemcpyAsync(ix->s_tmp.p, queries + b0 * ix->dim, nb * ix->dim * 4ull, cudaMemcpyDefault, s));
            ANNB_TRY(prepare_external(ix, ix->s_tmp.as<float>(), nb, &pq, s));
        } else {
            ANNB_TRY(prepare_self(ix, pos_begin + b0, nb, ivf, &pq, s));
        }
        ANNB_TRY(ix->s_ids.ensure(nb * k * 8ull));
        ANNB_TRY(ix->s_dist.ensure(nb * k * 4ull));
        ANNB_TRY(ix->s_cnt.ensure(nb * 4ull));
        if (ivf && mode == 1 && scatter) {
            // generate_knn: row of internal position p belongs to original id original_ids[p] (ivf.rs:476-486)
            std::vector<uint64_t> oid(nb);
            std::vector<uint64_t> ids(nb * k);
            std::vector<float> dist(out_dist ? nb * k : 0);
            std::vector<uint32_t> cnt(out_counts ? nb : 0);
            auto copy_out = [&]() -> int {
                ANNB_CUDA_CHECK(cudaMemcpyAsync(oid.data(), ix->d_original_ids + pos_begin + b0, nb * 8, cudaMemcpyDeviceToHost, s));
                ANNB_CUDA_CHECK(cudaMemcpyAsync(ids.data(), ix->s_ids.p, nb * k * 8ull, cudaMemcpyDeviceToHost, s));
                if (out_dist) ANNB_CUDA_CHECK(cudaMemcpyAsync(dist.data(), ix->s_dist.p, nb * k * 4ull, cudaMemcpyDeviceToHost, s));
                if (out_counts) ANNB_CUDA_CHECK(cudaMemcpyAsync(cnt.data(), ix->s_cnt.p, nb * 4ull, cudaMemcpyDeviceToHost, s));
                return ANNB_OK;
            };
            ANNB_TRY(run_batch(ix, ivf, pq, nb, k, nprobe, ix->s_ids.as<uint64_t>(), ix->s_dist.as<float>(), ix->s_cnt.as<uint32_t>(), s, true, copy_out));
            for (uint64_t i = 0; i < nb; i++) {
                std::memcpy(out_ids + oid[i] * k, ids.data() + i * k, k * 8ull);
                if (out_dist) std::memcpy(out_dist + oid[i] * k, dist.data() + i * k, k * 4ull);
                if (out_counts) out_counts[oid[i]] = cnt[i];
            }
        } else if (!ivf && mode == 1 && scatter == 2) {
            // kNN-graph rows: the batch was searched with k = graph degree + 1; drop every row's own id on the device
            const uint32_t kg = k - 1;
            ANNB_TRY(ix->s_route.ensure(nb * kg * 8ull));
            ANNB_TRY(ix->s_qcodes.ensure(nb * kg * 4ull));
            auto copy_out = [&]() -> int {
                knn_graph_rows_kernel<<<static_cast<uint32_t>(ceil_div<uint64_t>(nb, 128)), 128, 0, s>>>(ix->s_ids.as<uint64_t>(), ix->s_dist.as<float>(), nb, k,
                                                                                                   ix->id_base + pos_begin + b0, ix->s_route.as<uint64_t>(),
                                                                                                   ix->s_qcodes.as<float>(), ix->s_cnt.as<uint32_t>());
                ANNB_CUDA_CHECK(cudaGetLastError());
                ANNB_CUDA_CHECK(cudaMemcpyAsync(out_ids + b0 * kg, ix->s_route.p, nb * kg * 8ull, cudaMemcpyDefault, s));
                ANNB_CUDA_CHECK(cudaMemcpyAsync(out_dist + b0 * kg, ix->s_qcodes.p, nb * kg * 4ull, cudaMemcpyDefault, s));
                if (out_counts) ANNB_CUDA_CHECK(cudaMemcpyAsync(out_counts + b0, ix->s_cnt.p, nb * 4ull, cudaMemcpyDefault, s));
                return ANNB_OK;
            };
            ANNB_TRY(run_batch(ix, ivf, pq, nb, k, nprobe, ix->s_ids.as<uint64_t>(), ix->s_dist.as<float>(), ix->s_cnt.as<uint32_t>(), s, true, copy_out));
        } else {
            auto copy_out = [&]() -> int {
                ANNB_CUDA_CHECK(cudaMemcpyAsync(out_ids + b0 * k, ix->s_ids.p, nb * k * 8ull, cudaMemcpyDefault, s));
                if (out_dist) ANNB_CUDA_CHECK(cudaMemcpyAsync(out_dist + b0 * k, ix->s_dist.p, nb * k * 4ull, cudaMemcpyDefault, s));
                if (out_counts) ANNB_CUDA_CHECK(cudaMemcpyAsync(out_counts + b0, ix->s_cnt.p, nb * 4ull, cudaMemcpyDefault, s));
                return ANNB_OK;
            };
            ANNB_TRY(run_batch(ix, ivf, pq, nb, k, nprobe, ix->s_ids.as<uint64_t>(), ix->s_dist.as<float>(), ix->s_cnt.as<uint32_t>(), s, true, copy_out));
        }
    }
    return mark_call_done(ix, s);
}

static int common_create(annb_index* ix, int device) {
    ix->device = device;
    ANNB_CUDA_CHECK(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
    ANNB_CUDA_CHECK(cudaMallocHost(&ix->h_status, 64));
    return ANNB_OK;
}

}  // namespace annb

#include "multi.cuh"

using namespace annb;

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

const char* annb_last_error(void) { return g_last_error.c_str(); }
int annb_version(void) { return ANNB_VERSION_MAJOR * 1000 + ANNB_VERSION_MINOR; }

int annb_device_count(int* out) {
    if (!out) return fail(ANNB_ERR_INVALID_ARGUMENT, "null out");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        *out = 0;
        return fail(ANNB_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *out = c;
    return ANNB_OK;
}

int annb_parse_metric(const char* str) {
    if (!str) return -1;
    std::string s(str);
    for (auto& c : s) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
    if (s == "euclidean" || s == "l2") return ANNB_L2;
    if (s == "cosine") return ANNB_COSINE;
    if (s == "manhattan" || s == "l1") return ANNB_MANHATTAN;
    return -1;
}

void annb_destroy(annb_index* ix) {
    if (!ix) return;
    if (ix->multi) {   // front of a multi-device index: the shards own everything
        multi_destroy(ix->multi);
        delete ix;
        return;
    }
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(ix->device);
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    collect_timers(ix);
    tc_destroy(ix);
    tc_ivf_destroy(ix);
    tc_stream_destroy(ix);
    tc_coarse_destroy(ix);
    cudaFree(ix->d_rows); cudaFree(ix->d_norms); cudaFree(ix->d_norms_i); cudaFree(ix->d_scales);
    cudaFree(ix->d_centroids); cudaFree(ix->d_centroid_norms); cudaFree(ix->d_offsets); cudaFree(ix->d_original_ids); cudaFree(ix->d_list_order);
    for (DevBuf* b : {&ix->s_qpad, &ix->s_qcodes, &ix->s_route, &ix->s_cdist, &ix->s_probes, &ix->s_nprobes, &ix->s_keys, &ix->s_flags,
                      &ix->s_ids, &ix->s_dist, &ix->s_cnt, &ix->s_tmp, &ix->s_pairs, &ix->s_uncert, &ix->s_fbq, &ix->s_fbr, &ix->s_fbi, &ix->s_fbd, &ix->s_fbc})
        b->release();
    if (ix->stream) cudaStreamDestroy(ix->stream);
    if (ix->last_event) cudaEventDestroy(ix->last_event);
    if (ix->h_status) cudaFreeHost(ix->h_status);
    (void)cudaGetLastError();
    if (prev >= 0) cudaSetDevice(prev);
    delete ix;
}

// out[r][c] = src[r * rs + c * cs]: 32 x 32 tiles through shared memory, so a column-major source (rs == 1) is read and the
// row-major result written with coalesced accesses.
static __global__ void strided_to_rowmajor_kernel(const float* __restrict__ src, uint64_t nrows, uint32_t ncols, uint64_t rs, uint64_t cs,
                                                  float* __restrict__ out) {
    __shared__ float tile[32][33];
    const uint64_t r0 = static_cast<uint64_t>(blockIdx.x) * 32;
    const uint32_t c0 = blockIdx.y * 32;
    for (uint32_t j = threadIdx.y; j < 32; j += blockDim.y) {
        const uint64_t r = r0 + threadIdx.x;
        const uint32_t c = c0 + j;
        if (r < nrows && c < ncols) tile[j][threadIdx.x] = src[r * rs + c * cs];
    }
    __syncthreads();
    for (uint32_t j = threadIdx.y; j < 32; j += blockDim.y) {
        const uint64_t r = r0 + j;
        const uint32_t c = c0 + threadIdx.x;
        if (r < nrows && c < ncols) out[r * ncols + c] = tile[threadIdx.x][j];
    }
}

static bool is_device_pointer(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int annb_matrix_to_flat(const float* mat, uint64_t nrows, uint32_t ncols, int64_t row_stride, int64_t col_stride, float* out_rowmajor, int device) {
    if (!mat || !out_rowmajor || nrows == 0 || ncols == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "empty matrix");
    if (row_stride < 0 || col_stride < 0) return fail(ANNB_ERR_UNSUPPORTED, "negative strides (reversed views) are not supported");
    if ((row_stride == 0 && nrows > 1) || (col_stride == 0 && ncols > 1)) return fail(ANNB_ERR_INVALID_ARGUMENT, "zero stride");
    ANNB_DEVICE(device);
    const uint64_t rs = static_cast<uint64_t>(row_stride), cs = static_cast<uint64_t>(col_stride);
    const uint64_t total = nrows * ncols;
    if (cs == 1 && rs == ncols) {   // already contiguous row-major
        ANNB_CUDA_CHECK(cudaMemcpy(out_rowmajor, mat, total * 4, cudaMemcpyDefault));
        return ANNB_OK;
    }
    if (cs == 1) {                  // row-major with padded rows
        ANNB_CUDA_CHECK(cudaMemcpy2D(out_rowmajor, ncols * 4ull, mat, rs * 4, ncols * 4ull, nrows, cudaMemcpyDefault));
        return ANNB_OK;
    }
    const bool src_dev = is_device_pointer(mat), out_dev = is_device_pointer(out_rowmajor);
    DevBuf stage, result;
    struct Rel { DevBuf& a; DevBuf& b; ~Rel() { a.release(); b.release(); } } rel{stage, result};
    const float* d_src = mat;
    uint64_t k_rs = rs, k_cs = cs;
    if (!src_dev) {
        if (rs == 1) {              // column-major (faer's default layout), columns possibly padded: pack the columns on the way up
            ANNB_TRY(stage.ensure(total * 4));
            ANNB_CUDA_CHECK(cudaMemcpy2D(stage.p, nrows * 4, mat, cs * 4, nrows * 4, ncols, cudaMemcpyDefault));
            k_cs = nrows;
        } else {                    // both strides non-trivial: upload the spanned region as it is
            const uint64_t span = (nrows - 1) * rs + (ncols - 1) * cs + 1;
            if (span > 8 * total + 1024) return fail(ANNB_ERR_UNSUPPORTED, "matrix view is too sparse in its buffer (span > 8x the elements): copy it first");
            ANNB_TRY(stage.ensure(span * 4));
            ANNB_CUDA_CHECK(cudaMemcpy(stage.p, mat, span * 4, cudaMemcpyDefault));
        }
        d_src = stage.as<float>();
    }
    float* d_out = out_rowmajor;
    if (!out_dev) {
        ANNB_TRY(result.ensure(total * 4));
        d_out = result.as<float>();
    }
    if ((nrows + 31) / 32 > 0x7FFFFFFFull) return fail(ANNB_ERR_UNSUPPORTED, "too many rows");
    strided_to_rowmajor_kernel<<<dim3(static_cast<uint32_t>((nrows + 31) / 32), (ncols + 31) / 32), dim3(32, 8)>>>(d_src, nrows, ncols, k_rs, k_cs, d_out);
    ANNB_CUDA_CHECK(cudaGetLastError());
    if (!out_dev) ANNB_CUDA_CHECK(cudaMemcpy(out_rowmajor, d_out, total * 4, cudaMemcpyDefault));
    else ANNB_CUDA_CHECK(cudaDeviceSynchronize());
    return ANNB_OK;
}

// bincode 2 "standard" variable-length integers (the crate persists its Vec<usize> fields with them, src/serialise/mod.rs:51-64):
// u < 251 one byte; then a marker byte 251 / 252 / 253 followed by the value as little-endian u16 / u32 / u64.  Host code only.
int64_t annb_varint_encode_u64(const uint64_t* values, uint64_t count, uint8_t* out, uint64_t out_capacity) {
    if ((!values && count) || (!out && out_capacity)) { fail(ANNB_ERR_INVALID_ARGUMENT, "null buffer"); return -1; }
    uint64_t w = 0;
    for (uint64_t i = 0; i < count; i++) {
        const uint64_t v = values[i];
        const uint32_t nbytes = v < 251 ? 0u : (v <= 0xFFFFull ? 2u : (v <= 0xFFFFFFFFull ? 4u : 8u));
        if (w + 1 + nbytes > out_capacity) { fail(ANNB_ERR_INVALID_ARGUMENT, "varint output buffer too small"); return -1; }
        if (nbytes == 0) { out[w++] = static_cast<uint8_t>(v); continue; }
        out[w++] = nbytes == 2 ? 251 : (nbytes == 4 ? 252 : 253);
        for (uint32_t b = 0; b < nbytes; b++) out[w++] = static_cast<uint8_t>(v >> (8 * b));
    }
    return static_cast<int64_t>(w);
}
// Decodes `count` values; returns the number of bytes consumed, -1 if the buffer ends early or holds a marker this format
// does not produce for 64-bit values (254 = u128, 255 = reserved).
int64_t annb_varint_decode_u64(const uint8_t* buf, uint64_t len, uint64_t count, uint64_t* out) {
    if ((!buf && len) || (!out && count)) { fail(ANNB_ERR_INVALID_ARGUMENT, "null buffer"); return -1; }
    uint64_t r = 0;
    for (uint64_t i = 0; i < count; i++) {
        if (r >= len) return -1;
        const uint8_t m = buf[r++];
        if (m < 251) { out[i] = m; continue; }
        if (m > 253) return -1;
        const uint32_t nbytes = m == 251 ? 2u : (m == 252 ? 4u : 8u);
        if (r + nbytes > len) return -1;
        uint64_t v = 0;
        for (uint32_t b = 0; b < nbytes; b++) v |= static_cast<uint64_t>(buf[r + b]) << (8 * b);
        r += nbytes;
        out[i] = v;
    }
    return static_cast<int64_t>(r);
}

int annb_flat_create(annb_index** out, const float* data, uint64_t n, uint32_t dim, int dtype, int metric,
                     const float* sq8_scales, uint64_t id_base, int device) {
    if (!out) return fail(ANNB_ERR_INVALID_ARGUMENT, "null out");
    *out = nullptr;
    if (metric == ANNB_MANHATTAN) return fail(ANNB_ERR_DISTANCE_NOT_SUPPORTED, "Manhattan distance is not supported by the GPU / quantised indices");
    if (metric != ANNB_L2 && metric != ANNB_COSINE) return fail(ANNB_ERR_INVALID_ARGUMENT, "unknown metric");
    if (dtype < ANNB_F32 || dtype > ANNB_SQ8) return fail(ANNB_ERR_INVALID_ARGUMENT, "unknown dtype");
    if (!data || n == 0 || dim == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "empty data");
    if (n >= 0xFFFFFFFFull) return fail(ANNB_ERR_UNSUPPORTED, "n must be < 2^32 - 1 per handle (shard larger indices)");
    ANNB_DEVICE(device);
    annb_index* ix = new annb_index();
    struct Cleanup { annb_index*& p; bool armed = true; ~Cleanup() { if (armed) annb_destroy(p); } } cleanup{ix};
    ANNB_TRY(common_create(ix, device));
    ix->dtype = dtype; ix->metric = metric; ix->n = n; ix->n_total = n; ix->dim = dim; ix->id_base = id_base;
    ix->row_bytes = padded_row_bytes(dim, dtype);
    cudaStream_t s = ix->stream;

    // stage the f32 matrix, padded to 16-byte rows
    const uint32_t ld = round_up(dim * 4u, 16u) / 4u;
    float* d_f32 = nullptr;
    DevBuf raw;
    ANNB_TRY(dmalloc(&d_f32, n * ld, dtype == ANNB_F32 ? ix : nullptr));
    struct FreeTmp { float*& p; bool keep; ~FreeTmp() { if (!keep && p) cudaFree(p); } } free_tmp{d_f32, dtype == ANNB_F32};
    if (ld == dim) {
        ANNB_CUDA_CHECK(cudaMemcpyAsync(d_f32, data, n * dim * 4ull, cudaMemcpyDefault, s));
    } else {
        ANNB_CUDA_CHECK(cudaMemsetAsync(d_f32, 0, n * ld * 4ull, s));
        ANNB_CUDA_CHECK(cudaMemcpy2DAsync(d_f32, ld * 4ull, data, dim * 4ull, dim * 4ull, n, cudaMemcpyDefault, s));
    }
    if (dtype == ANNB_F32 || dtype == ANNB_BF16) {
        if (metric == ANNB_COSINE) {  // norms of the un-rounded rows (exhaustive.rs:86-96, exhaustive_bf16.rs:101-111)
            ANNB_TRY(dmalloc(&ix->d_norms, n, ix));
            row_norms_f32_kernel<<<grid_for(n, 128, 1u << 30), 128, 0, s>>>(d_f32, ld, dim, n, ix->d_norms, 0);
            ANNB_CUDA_CHECK(cudaGetLastError());
        }
        if (dtype == ANNB_F32) {
            ix->d_rows = reinterpret_cast<uint8_t*>(d_f32);
        } else {
            uint16_t* d_b = nullptr;
            ANNB_TRY(dmalloc(&d_b, n * (ix->row_bytes / 2), ix));
            ix->d_rows = reinterpret_cast<uint8_t*>(d_b);
            encode_bf16_kernel<<<grid_for(n * (ix->row_bytes / 2), 256), 256, 0, s>>>(d_f32, ld, dim, d_b, ix->row_bytes / 2, n);
            ANNB_CUDA_CHECK(cudaGetLastError());
        }
    } else {
        // ExhaustiveSq8Index::new (exhaustive_sq8.rs:104-151)
        if (metric == ANNB_COSINE) {
            normalise_rows_f32_kernel<<<grid_for(n, 128, 1u << 30), 128, 0, s>>>(d_f32, ld, dim, n);
            ANNB_CUDA_CHECK(cudaGetLastError());
        }
        ANNB_TRY(dmalloc(&ix->d_scales, dim, ix));
        if (sq8_scales) {
            ANNB_CUDA_CHECK(cudaMemcpyAsync(ix->d_scales, sq8_scales, dim * 4ull, cudaMemcpyDefault, s));
        } else {
            uint32_t* d_max = nullptr;
            ANNB_TRY(dmalloc(&d_max, dim, nullptr));
            ANNB_CUDA_CHECK(cudaMemsetAsync(d_max, 0, dim * 4ull, s));
            dim3 g(static_cast<uint32_t>(std::min<uint64_t>(n, 2048)), ceil_div(dim, 128u));
            sq8_absmax_kernel<<<g, 128, 0, s>>>(d_f32, ld, dim, n, d_max);
            sq8_scales_kernel<<<ceil_div(dim, 128u), 128, 0, s>>>(d_max, dim, ix->d_scales);
            cudaError_t e = cudaGetLastError();
            cudaStreamSynchronize(s);
            cudaFree(d_max);
            ANNB_CUDA_CHECK(e);
        }
        int8_t* d_c = nullptr;
        ANNB_TRY(dmalloc(&d_c, n * ix->row_bytes, ix));
        ix->d_rows = reinterpret_cast<uint8_t*>(d_c);
        sq8_encode_kernel<<<grid_for(n * ix->row_bytes, 256), 256, 0, s>>>(d_f32, ld, dim, ix->d_scales, d_c, ix->row_bytes, n);
        ANNB_CUDA_CHECK(cudaGetLastError());
        if (metric == ANNB_COSINE) {
            ANNB_TRY(dmalloc(&ix->d_norms_i, n, ix));
            sq8_row_norms_kernel<<<grid_for(n, 128, 1u << 30), 128, 0, s>>>(d_c, ix->row_bytes, dim, n, ix->d_norms_i);
            ANNB_CUDA_CHECK(cudaGetLastError());
        }
    }
    ANNB_CUDA_CHECK(cudaStreamSynchronize(s));
    ANNB_TRY(tc_flat_prepare(ix));
    ANNB_CUDA_CHECK(cudaStreamSynchronize(s));
    cleanup.armed = false;
    *out = ix;
    return ANNB_OK;
}

int annb_flat_search(const annb_index* index, const float* queries, uint64_t nq, uint32_t dim, uint32_t k,
                     uint64_t* out_ids, float* out_dist, uint32_t* out_counts) {
    if (index && dim != index->dim)
        return fail(ANNB_ERR_DIMENSION_MISMATCH, "query dim " + std::to_string(dim) + " != index dim " + std::to_string(index->dim));
    if (index && index->multi) return multi_search(const_cast<annb_index*>(index), false, 0, queries, 0, nq, k, 0, out_ids, out_dist, out_counts);
    return search_host(const_cast<annb_index*>(index), false, 0, queries, 0, nq, k, 0, 0, out_ids, out_dist, out_counts);
}

int annb_flat_search_self(const annb_index* index, uint64_t row_begin, uint64_t row_end, uint32_t k,
                          uint64_t* out_ids, float* out_dist, uint32_t* out_counts) {
    if (!index) return fail(ANNB_ERR_INVALID_ARGUMENT, "null index");
    if (row_begin > row_end || row_end > index->n) return fail(ANNB_ERR_INVALID_ARGUMENT, "row range outside the index");
    if (index->multi) return multi_search(const_cast<annb_index*>(index), false, 1, nullptr, row_begin, row_end - row_begin, k, 0, out_ids, out_dist, out_counts);
    return search_host(const_cast<annb_index*>(index), false, 1, nullptr, row_begin, row_end - row_begin, k, 0, 0, out_ids, out_dist, out_counts);
}

int annb_flat_knn_graph(const annb_index* index, uint64_t row_begin, uint64_t row_end, uint32_t k, uint64_t* out_pid, float* out_dist, uint32_t* out_counts) {
    if (!index) return fail(ANNB_ERR_INVALID_ARGUMENT, "null index");
    if (index->is_ivf) return fail(ANNB_ERR_INVALID_ARGUMENT, "not a flat index");
    if (!out_pid || !out_dist || k == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "null buffer / k == 0");
    if (row_begin > row_end || row_end > index->n) return fail(ANNB_ERR_INVALID_ARGUMENT, "row range outside the index");
    if (index->dtype != ANNB_F32) return fail(ANNB_ERR_UNSUPPORTED, "the kNN-graph hand-off is defined for f32 indices (KnnGraphGpu<T> carries the vectors themselves)");
    const uint64_t nq = row_end - row_begin;
    if (!index->multi) return search_host(const_cast<annb_index*>(index), false, 1, nullptr, row_begin, nq, k + 1, 0, 2, out_pid, out_dist, out_counts);
    // multi-device handle: the merged [nq][k + 1] rows arrive on the host; the (cheap, integer) self filter runs here
    std::vector<uint64_t> ids(nq * (k + 1ull));
    std::vector<float> dist(nq * (k + 1ull));
    ANNB_TRY(multi_search(const_cast<annb_index*>(index), false, 1, nullptr, row_begin, nq, k + 1, 0, ids.data(), dist.data(), nullptr));
    for (uint64_t q = 0; q < nq; q++) {
        uint32_t w = 0;
        for (uint32_t j = 0; j <= k && w < k; j++) {
            const uint64_t id = ids[q * (k + 1ull) + j];
            if (id == 0xFFFFFFFFFFFFFFFFull || id == row_begin + q) continue;
            out_pid[q * k + w] = id;
            out_dist[q * k + w] = dist[q * (k + 1ull) + j];
            w++;
        }
        if (out_counts) out_counts[q] = w;
        for (; w < k; w++) { out_pid[q * k + w] = 0x7FFFFFFFull; out_dist[q * k + w] = 3.402823466e+38f; }
    }
    return ANNB_OK;
}

int annb_flat_search_dev(const annb_index* index, const float* d_queries, uint64_t nq, uint32_t dim, uint32_t k,
                         uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts, void* stream) {
    annb_index* ix = const_cast<annb_index*>(index);
    if (!ix || ix->is_ivf) return fail(ANNB_ERR_INVALID_ARGUMENT, "not a flat index");
    if (dim != ix->dim) return fail(ANNB_ERR_DIMENSION_MISMATCH, "query dim " + std::to_string(dim) + " != index dim " + std::to_string(ix->dim));
    if (!d_queries || !d_out_ids || k == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "null buffer / k == 0");
    if (ix->multi) {   // buffers live on the first device; the call is synchronous (the shards run on their own streams)
        ANNB_DEVICE(ix->device);
        ANNB_CUDA_CHECK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
        return multi_search(ix, false, 0, d_queries, 0, nq, k, 0, d_out_ids, d_out_dist, d_out_counts);
    }
    ANNB_DEVICE(ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ANNB_TRY(order_after_previous(ix, s));
    for (uint64_t b0 = 0; b0 < nq; b0 += QUERY_BATCH) {
        const uint64_t nb = std::min<uint64_t>(QUERY_BATCH, nq - b0);
        PreparedQueries pq;
        ANNB_TRY(prepare_external(ix, d_queries + b0 * dim, nb, &pq, s));
        ANNB_TRY(run_batch(ix, false, pq, nb, k, 0, d_out_ids + b0 * k, d_out_dist ? d_out_dist + b0 * k : nullptr,
                           d_out_counts ? d_out_counts + b0 : nullptr, s, ix->opt_async_dev == 0, []() -> int { return ANNB_OK; }));
    }
    return mark_call_done(ix, s);
}

// out[list[i]] = src[i]
static __global__ void scatter_u32_kernel(const uint32_t* __restrict__ list, uint32_t count, const uint32_t* __restrict__ src, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[list[i]] = src[i];
}

// The exact CUDA-core assignment kernel (direct_assign arithmetic, k_means_utils.rs:2119-2195) on rows [0, nr) of d_x.
static int assign_simt(const float* d_x, uint32_t ld, uint64_t nr, const float* d_c, const float* d_aux, uint32_t nlist, uint32_t dim, bool cosine,
                       uint32_t* d_a) {
    TileParams p{};
    p.rows = reinterpret_cast<const uint8_t*>(d_c); p.n_rows = nlist; p.row_bytes = ld * 4; p.row_aux = d_aux;
    p.queries = reinterpret_cast<const uint8_t*>(d_x); p.q_bytes = ld * 4; p.nq = nr; p.dim = dim;
    p.assign_out = d_a; p.assign_cosine = cosine;
    const size_t smem = tile_kernel_smem(p.row_bytes, p.q_bytes, 0, false);
    return launch_tile<0, QT_F32, MET_DOT, EPI_ARGMAX>(p, dim3(static_cast<uint32_t>(ceil_div<uint64_t>(nr, CTA_QUERIES)), 1), smem, 0);
}

// One chunk of rows: tensor-core pre-selection + exact scores + certificate (flat_tc.cu), the rows that fail it redone on
// the exact kernel; without a tensor state (small tables, wide rows, ANNB200_ASSIGN_PATH=simt) the exact kernel alone.
struct AssignScratch {
    DevBuf uncert, fb_rows, fb_assign;
    uint64_t redone = 0;
    ~AssignScratch() { uncert.release(); fb_rows.release(); fb_assign.release(); }
};
static int assign_chunk(TcAssignState* tcs, AssignScratch& sc, const float* d_x, uint32_t ld, uint64_t nr, const float* d_c, const float* d_aux,
                        uint32_t nlist, uint32_t dim, bool cosine, uint32_t* d_a) {
    if (!tcs) return assign_simt(d_x, ld, nr, d_c, d_aux, nlist, dim, cosine, d_a);
    ANNB_TRY(sc.uncert.ensure((nr + 1) * 4));
    ANNB_TRY(tc_assign_run(tcs, d_x, ld, nr, d_c, ld, d_aux, cosine, d_a, sc.uncert.as<uint32_t>(), 0));
    uint32_t n_unc = 0;
    ANNB_CUDA_CHECK(cudaMemcpy(&n_unc, sc.uncert.p, 4, cudaMemcpyDeviceToHost));
    if (n_unc == 0) return ANNB_OK;
    sc.redone += n_unc;
    const uint32_t* list = sc.uncert.as<uint32_t>() + 1;
    ANNB_TRY(sc.fb_rows.ensure(static_cast<size_t>(n_unc) * ld * 4));
    ANNB_TRY(sc.fb_assign.ensure(static_cast<size_t>(n_unc) * 4));
    gather_rows_kernel<<<grid_for(static_cast<uint64_t>(n_unc) * (ld >> 2), 256), 256>>>(reinterpret_cast<const uint8_t*>(d_x), ld * 4, list, n_unc, sc.fb_rows.as<uint8_t>());
    ANNB_CUDA_CHECK(cudaGetLastError());
    ANNB_TRY(assign_simt(sc.fb_rows.as<float>(), ld, n_unc, d_c, d_aux, nlist, dim, cosine, sc.fb_assign.as<uint32_t>()));
    scatter_u32_kernel<<<ceil_div(n_unc, 256u), 256>>>(list, n_unc, sc.fb_assign.as<uint32_t>(), d_a);
    ANNB_CUDA_CHECK(cudaGetLastError());
    return ANNB_OK;
}
static bool assign_tensor_enabled() {
    const char* e = std::getenv("ANNB200_ASSIGN_PATH");   // debugging / A-B switch: "simt" keeps the build-side assignment on the CUDA cores
    return !(e && std::string(e) == "simt");
}
static thread_local uint64_t g_assign_redone = 0;   // rows of the last assign / Lloyd call that failed the certificate

int annb_ivf_assign(const float* data, uint64_t n, uint32_t dim, const float* centroids, const float* centroid_norms,
                    uint32_t nlist, int metric, uint32_t* out_assign, int device) {
    if (metric == ANNB_MANHATTAN) return fail(ANNB_ERR_DISTANCE_NOT_SUPPORTED, "Manhattan distance is not supported");
    if (!data || !centroids || !out_assign || n == 0 || dim == 0 || nlist == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "empty input");
    ANNB_DEVICE(device);
    const uint32_t ld = round_up(dim * 4u, 16u) / 4u;
    float *d_c = nullptr, *d_aux = nullptr, *d_x = nullptr, *d_cn = nullptr;
    uint32_t* d_a = nullptr;
    const uint64_t chunk = std::min<uint64_t>(n, 1ull << 20);
    struct Free { float*& a; float*& b; float*& c; float*& e; uint32_t*& d; ~Free() { cudaFree(a); cudaFree(b); cudaFree(c); cudaFree(e); cudaFree(d); } } fr{d_c, d_aux, d_x, d_cn, d_a};
    ANNB_TRY(dmalloc(&d_c, static_cast<size_t>(nlist) * ld, nullptr));
    ANNB_TRY(dmalloc(&d_aux, nlist, nullptr));
    ANNB_TRY(dmalloc(&d_x, chunk * ld, nullptr));
    ANNB_TRY(dmalloc(&d_a, chunk, nullptr));
    ANNB_CUDA_CHECK(cudaMemset(d_c, 0, static_cast<size_t>(nlist) * ld * 4));
    ANNB_CUDA_CHECK(cudaMemcpy2D(d_c, ld * 4ull, centroids, dim * 4ull, dim * 4ull, nlist, cudaMemcpyDefault));
    if (centroid_norms && metric == ANNB_COSINE) {
        ANNB_TRY(dmalloc(&d_cn, nlist, nullptr));
        ANNB_CUDA_CHECK(cudaMemcpy(d_cn, centroid_norms, nlist * 4ull, cudaMemcpyDefault));
    }
    assign_aux_kernel<<<ceil_div(nlist, 128u), 128>>>(d_c, ld, dim, nlist, metric == ANNB_COSINE, d_cn, d_aux);
    ANNB_CUDA_CHECK(cudaGetLastError());
    TcAssignState* tcs = nullptr;
    struct FreeTc { TcAssignState*& p; ~FreeTc() { tc_assign_destroy(p); } } ftc{tcs};
    if (assign_tensor_enabled() && n >= 4096) ANNB_TRY(tc_assign_create(&tcs, dim, nlist));
    if (tcs) ANNB_TRY(tc_assign_set_centroids(tcs, d_c, ld, d_aux, metric == ANNB_COSINE, 0));
    AssignScratch sc;
    for (uint64_t r0 = 0; r0 < n; r0 += chunk) {
        const uint64_t nr = std::min<uint64_t>(chunk, n - r0);
        if (ld != dim) ANNB_CUDA_CHECK(cudaMemset(d_x, 0, nr * ld * 4ull));
        ANNB_CUDA_CHECK(cudaMemcpy2D(d_x, ld * 4ull, data + r0 * dim, dim * 4ull, dim * 4ull, nr, cudaMemcpyDefault));
        ANNB_TRY(assign_chunk(tcs, sc, d_x, ld, nr, d_c, d_aux, nlist, dim, metric == ANNB_COSINE, d_a));
        ANNB_CUDA_CHECK(cudaMemcpy(out_assign + r0, d_a, nr * 4ull, cudaMemcpyDefault));
    }
    g_assign_redone = sc.redone;
    return ANNB_OK;
}

/* Diagnostic: rows of this thread's last annb_ivf_assign / annb_kmeans_lloyd call that failed the tensor path's certificate
 * and were redone on the exact kernel (0 when the call ran on the exact kernel throughout). */
uint64_t annb_assign_last_redone(void) { return g_assign_redone; }

// centroid c <- (centroid c * w + data row j) / (w + 1), separate multiply / add / divide as the reference's f32 loop
// (adjust_centers, src/utils/k_means_utils.rs:1020-1025); one block per moved centroid
static __global__ void kmeans_adjust_kernel(const uint32_t* __restrict__ moves, float* __restrict__ cent, uint32_t ld, uint32_t dim, const float* __restrict__ x) {
    const uint32_t c = moves[3 * blockIdx.x], j = moves[3 * blockIdx.x + 1];
    const float w = static_cast<float>(moves[3 * blockIdx.x + 2]);
    const float denom = __fadd_rn(w, 1.0f);
    for (uint32_t d = threadIdx.x; d < dim; d += blockDim.x) {
        float* cc = cent + static_cast<uint64_t>(c) * ld + d;
        *cc = __fdiv_rn(__fadd_rn(__fmul_rn(*cc, w), x[static_cast<uint64_t>(j) * ld + d]), denom);
    }
}

int annb_kmeans_lloyd(const float* data, uint64_t n, uint32_t dim, float* centroids, uint32_t nlist, int metric, uint32_t max_iters,
                      uint32_t* out_iters, int device) {
    return annb_kmeans_lloyd_balanced(data, n, dim, centroids, nlist, metric, max_iters, 0, 0, out_iters, nullptr, device);
}

int annb_kmeans_lloyd_balanced(const float* data, uint64_t n, uint32_t dim, float* centroids, uint32_t nlist, int metric, uint32_t max_iters,
                               int balanced, uint64_t seed, uint32_t* out_iters, uint64_t* out_adjusted, int device) {
    if (metric == ANNB_MANHATTAN) return fail(ANNB_ERR_DISTANCE_NOT_SUPPORTED, "Manhattan distance is not supported");
    if (metric != ANNB_L2 && metric != ANNB_COSINE) return fail(ANNB_ERR_INVALID_ARGUMENT, "unknown metric");
    if (!data || !centroids || n == 0 || dim == 0 || nlist == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "empty input");
    if (nlist > n) return fail(ANNB_ERR_TOO_FEW_SAMPLES, std::to_string(n) + " training samples for " + std::to_string(nlist) + " centroids");
    ANNB_DEVICE(device);
    const uint32_t ld = round_up(dim * 4u, 16u) / 4u;
    float *d_c = nullptr, *d_aux = nullptr, *d_x = nullptr, *d_cn = nullptr;
    uint32_t *d_a = nullptr, *d_prev = nullptr, *d_cnt = nullptr;
    double* d_sums = nullptr;
    unsigned long long* d_changed = nullptr;
    struct Free {
        std::vector<void**> p;
        ~Free() { for (void** q : p) cudaFree(*q); }
    } fr{{reinterpret_cast<void**>(&d_c), reinterpret_cast<void**>(&d_aux), reinterpret_cast<void**>(&d_x), reinterpret_cast<void**>(&d_cn),
          reinterpret_cast<void**>(&d_a), reinterpret_cast<void**>(&d_prev), reinterpret_cast<void**>(&d_cnt), reinterpret_cast<void**>(&d_sums),
          reinterpret_cast<void**>(&d_changed)}};
    ANNB_TRY(dmalloc(&d_c, static_cast<size_t>(nlist) * ld, nullptr));
    ANNB_TRY(dmalloc(&d_aux, nlist, nullptr));
    ANNB_TRY(dmalloc(&d_cn, nlist, nullptr));
    ANNB_TRY(dmalloc(&d_x, n * ld, nullptr));
    ANNB_TRY(dmalloc(&d_a, n, nullptr));
    ANNB_TRY(dmalloc(&d_prev, n, nullptr));
    ANNB_TRY(dmalloc(&d_cnt, nlist, nullptr));
    ANNB_TRY(dmalloc(&d_sums, static_cast<size_t>(nlist) * dim, nullptr));
    ANNB_TRY(dmalloc(&d_changed, 1, nullptr));
    ANNB_CUDA_CHECK(cudaMemset(d_c, 0, static_cast<size_t>(nlist) * ld * 4));
    ANNB_CUDA_CHECK(cudaMemcpy2D(d_c, ld * 4ull, centroids, dim * 4ull, dim * 4ull, nlist, cudaMemcpyDefault));
    if (ld != dim) ANNB_CUDA_CHECK(cudaMemset(d_x, 0, n * ld * 4ull));
    ANNB_CUDA_CHECK(cudaMemcpy2D(d_x, ld * 4ull, data, dim * 4ull, dim * 4ull, n, cudaMemcpyDefault));
    ANNB_CUDA_CHECK(cudaMemset(d_prev, 0xFF, n * 4ull));      // usize::MAX: every point counts as changed in the first iteration
    const uint64_t change_floor = std::max<uint64_t>(n / 10000, 1);
    TcAssignState* tcs = nullptr;
    struct FreeTc { TcAssignState*& p; ~FreeTc() { tc_assign_destroy(p); } } ftc{tcs};
    if (assign_tensor_enabled() && n >= 4096) ANNB_TRY(tc_assign_create(&tcs, dim, nlist));
    AssignScratch sc;
    // balancing (adjust_centers): the donor walk is a short serial pass over the assignments -- it runs on the host between
    // two device steps; the centroid moves themselves happen on the device
    std::vector<uint32_t> h_assign, h_cnt, h_moves;
    DevBuf d_moves;
    struct RelMoves { DevBuf& b; ~RelMoves() { b.release(); } } rel_moves{d_moves};
    uint64_t last_adjusted = 0, total_adjusted = 0;
    if (n >= 0xFFFFFFFFull) return fail(ANNB_ERR_UNSUPPORTED, "n must be < 2^32 - 1");
    uint32_t it = 0;
    for (; it < max_iters; it++) {
        // assignment: direct_assign arithmetic (k_means_utils.rs:2119-2195) on the current centroids
        if (metric == ANNB_COSINE) row_norms_f32_kernel<<<ceil_div(nlist, 128u), 128>>>(d_c, ld, dim, nlist, d_cn, 0);   // calculate_l2_norm
        assign_aux_kernel<<<ceil_div(nlist, 128u), 128>>>(d_c, ld, dim, nlist, metric == ANNB_COSINE, d_cn, d_aux);
        ANNB_CUDA_CHECK(cudaGetLastError());
        if (tcs) ANNB_TRY(tc_assign_set_centroids(tcs, d_c, ld, d_aux, metric == ANNB_COSINE, 0));
        const uint64_t chunk = 1ull << 20;
        for (uint64_t r0 = 0; r0 < n; r0 += chunk) {
            const uint64_t nr = std::min<uint64_t>(chunk, n - r0);
            ANNB_TRY(assign_chunk(tcs, sc, d_x + r0 * ld, ld, nr, d_c, d_aux, nlist, dim, metric == ANNB_COSINE, d_a + r0));
        }
        // convergence is tested before the update (k_means_utils.rs:1611-1624)
        ANNB_CUDA_CHECK(cudaMemset(d_changed, 0, 8));
        kmeans_changed_kernel<<<static_cast<uint32_t>(ceil_div<uint64_t>(n, 256)), 256>>>(d_a, d_prev, n, d_changed);
        unsigned long long changed = 0;
        ANNB_CUDA_CHECK(cudaMemcpy(&changed, d_changed, 8, cudaMemcpyDeviceToHost));
        if (changed <= change_floor && last_adjusted == 0) break;   // (the reference stays in until balancing has nothing left to do, :1618)
        ANNB_CUDA_CHECK(cudaMemset(d_sums, 0, static_cast<size_t>(nlist) * dim * 8));
        ANNB_CUDA_CHECK(cudaMemset(d_cnt, 0, nlist * 4ull));
        kmeans_accumulate_kernel<<<grid_for(n * dim, 256, 148u * 32u), 256>>>(d_x, ld, dim, n, d_a, d_sums, d_cnt);
        kmeans_update_kernel<<<static_cast<uint32_t>(ceil_div<uint64_t>(static_cast<uint64_t>(nlist) * dim, 256)), 256>>>(d_sums, d_cnt, d_c, ld, dim, nlist);
        ANNB_CUDA_CHECK(cudaGetLastError());
        if (balanced) {
            // adjust_centers(seed + iter) (src/utils/k_means_utils.rs:979-1030, hook at :1668-1682)
            h_assign.resize(n);
            h_cnt.resize(nlist);
            ANNB_CUDA_CHECK(cudaMemcpy(h_assign.data(), d_a, n * 4ull, cudaMemcpyDeviceToHost));
            ANNB_CUDA_CHECK(cudaMemcpy(h_cnt.data(), d_cnt, nlist * 4ull, cudaMemcpyDeviceToHost));
            h_moves.clear();
            const double average = static_cast<double>(n) / static_cast<double>(nlist), floor_ = average * 0.25;
            uint64_t cursor = (seed + it) % n;
            for (uint32_t c = 0; c < nlist; c++) {
                if (static_cast<double>(h_cnt[c]) > floor_) continue;
                int64_t donor = -1;
                for (uint64_t t = 0; t < n; t++) {
                    cursor = (cursor + 715827883ull) % n;
                    const uint32_t owner = h_assign[cursor];
                    if (owner != c && static_cast<double>(h_cnt[owner]) > average) { donor = static_cast<int64_t>(cursor); break; }
                }
                if (donor < 0) continue;
                h_moves.push_back(c);
                h_moves.push_back(static_cast<uint32_t>(donor));
                h_moves.push_back(std::min<uint32_t>(h_cnt[c], 5u));
            }
            last_adjusted = h_moves.size() / 3;
            total_adjusted += last_adjusted;
            if (last_adjusted) {
                ANNB_TRY(d_moves.ensure(h_moves.size() * 4ull));
                ANNB_CUDA_CHECK(cudaMemcpy(d_moves.p, h_moves.data(), h_moves.size() * 4ull, cudaMemcpyHostToDevice));
                kmeans_adjust_kernel<<<static_cast<uint32_t>(last_adjusted), 128>>>(d_moves.as<uint32_t>(), d_c, ld, dim, d_x);
                ANNB_CUDA_CHECK(cudaGetLastError());
            }
        }
    }
    ANNB_CUDA_CHECK(cudaMemcpy2D(centroids, dim * 4ull, d_c, ld * 4ull, dim * 4ull, nlist, cudaMemcpyDefault));
    g_assign_redone = sc.redone;
    if (out_iters) *out_iters = it;
    if (out_adjusted) *out_adjusted = total_adjusted;
    return ANNB_OK;
}

int annb_ivf_create(annb_index** out, const void* vectors, const void* norms, const float* centroids,
                    const float* centroid_norms, const uint64_t* offsets, const uint64_t* original_ids,
                    uint64_t n, uint32_t dim, uint32_t nlist, int dtype, int metric, const float* sq8_scales,
                    uint32_t list_begin, uint32_t list_end, int device) {
    if (!out) return fail(ANNB_ERR_INVALID_ARGUMENT, "null out");
    *out = nullptr;
    if (metric == ANNB_MANHATTAN) return fail(ANNB_ERR_DISTANCE_NOT_SUPPORTED, "Manhattan distance is not supported by the IVF indices");
    if (metric != ANNB_L2 && metric != ANNB_COSINE) return fail(ANNB_ERR_INVALID_ARGUMENT, "unknown metric");
    if (dtype < ANNB_F32 || dtype > ANNB_SQ8) return fail(ANNB_ERR_INVALID_ARGUMENT, "unknown dtype");
    if (!vectors || !centroids || !offsets || !original_ids || n == 0 || dim == 0 || nlist == 0)
        return fail(ANNB_ERR_INVALID_ARGUMENT, "empty index contents");
    if (list_begin > list_end || list_end > nlist) return fail(ANNB_ERR_INVALID_ARGUMENT, "bad list range");
    if (dtype == ANNB_SQ8 && !sq8_scales) return fail(ANNB_ERR_INVALID_ARGUMENT, "SQ8 index needs its codebook scales");
    if (metric == ANNB_COSINE && !norms) return fail(ANNB_ERR_INVALID_ARGUMENT, "cosine index needs per-vector norms");
    if (metric == ANNB_COSINE && dtype != ANNB_SQ8 && !centroid_norms) return fail(ANNB_ERR_INVALID_ARGUMENT, "cosine index needs centroid norms");
    std::vector<uint64_t> h_off(nlist + 1);
    {
        cudaError_t e = cudaMemcpy(h_off.data(), offsets, (nlist + 1) * 8ull, cudaMemcpyDefault);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(ANNB_ERR_CUDA, std::string("offsets copy: ") + cudaGetErrorString(e) + " (libannb200 has no CPU fallback)"); }
    }
    for (uint32_t c = 0; c < nlist; c++)
        if (h_off[c] > h_off[c + 1]) return fail(ANNB_ERR_INVALID_ARGUMENT, "offsets must be non-decreasing");
    if (h_off[nlist] != n) return fail(ANNB_ERR_INVALID_ARGUMENT, "offsets[nlist] must equal n");
    const uint64_t n_local = h_off[list_end] - h_off[list_begin];
    if (n >= 0xFFFFFFFFull) return fail(ANNB_ERR_UNSUPPORTED, "n must be < 2^32 - 1");
    ANNB_DEVICE(device);
    annb_index* ix = new annb_index();
    struct Cleanup { annb_index*& p; bool armed = true; ~Cleanup() { if (armed) annb_destroy(p); } } cleanup{ix};
    ANNB_TRY(common_create(ix, device));
    ix->is_ivf = true; ix->dtype = dtype; ix->metric = metric; ix->n = n_local; ix->n_total = n; ix->dim = dim;
    ix->nlist = nlist; ix->list_begin = list_begin; ix->list_end = list_end; ix->shard_row0 = h_off[list_begin];
    ix->h_offsets = h_off;
    ix->row_bytes = padded_row_bytes(dim, dtype);
    cudaStream_t s = ix->stream;
    const uint32_t src_row = dim * elem_bytes(dtype);
    ANNB_TRY(dmalloc(&ix->d_rows, std::max<uint64_t>(n_local, 1) * ix->row_bytes, ix));
    if (n_local) {
        if (src_row == ix->row_bytes) {
            ANNB_CUDA_CHECK(cudaMemcpyAsync(ix->d_rows, vectors, n_local * src_row, cudaMemcpyDefault, s));
        } else {
            ANNB_CUDA_CHECK(cudaMemsetAsync(ix->d_rows, 0, n_local * ix->row_bytes, s));
            ANNB_CUDA_CHECK(cudaMemcpy2DAsync(ix->d_rows, ix->row_bytes, vectors, src_row, src_row, n_local, cudaMemcpyDefault, s));
        }
    }
    if (metric == ANNB_COSINE && n_local) {
        if (dtype == ANNB_SQ8) {
            ANNB_TRY(dmalloc(&ix->d_norms_i, n_local, ix));
            ANNB_CUDA_CHECK(cudaMemcpyAsync(ix->d_norms_i, norms, n_local * 4ull, cudaMemcpyDefault, s));
        } else {
            ANNB_TRY(dmalloc(&ix->d_norms, n_local, ix));
            ANNB_CUDA_CHECK(cudaMemcpyAsync(ix->d_norms, norms, n_local * 4ull, cudaMemcpyDefault, s));
        }
    }
    ix->cent_ld = round_up(dim * 4u, 16u) / 4u;
    ANNB_TRY(dmalloc(&ix->d_centroids, static_cast<size_t>(nlist) * ix->cent_ld, ix));
    ANNB_CUDA_CHECK(cudaMemsetAsync(ix->d_centroids, 0, static_cast<size_t>(nlist) * ix->cent_ld * 4, s));
    ANNB_CUDA_CHECK(cudaMemcpy2DAsync(ix->d_centroids, ix->cent_ld * 4ull, centroids, dim * 4ull, dim * 4ull, nlist, cudaMemcpyDefault, s));
    if (metric == ANNB_COSINE && dtype != ANNB_SQ8) {
        ANNB_TRY(dmalloc(&ix->d_centroid_norms, nlist, ix));
        ANNB_CUDA_CHECK(cudaMemcpyAsync(ix->d_centroid_norms, centroid_norms, nlist * 4ull, cudaMemcpyDefault, s));
    }
    ANNB_TRY(dmalloc(&ix->d_offsets, nlist + 1, ix));
    ANNB_CUDA_CHECK(cudaMemcpyAsync(ix->d_offsets, h_off.data(), (nlist + 1) * 8ull, cudaMemcpyHostToDevice, s));
    {   // lists by descending length: the tensor-core scan hands out its (list x query group) tasks longest first, so the
        // tail of the dynamic schedule is made of the shortest lists
        std::vector<uint32_t> order(nlist);
        for (uint32_t c = 0; c < nlist; c++) order[c] = c;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return h_off[a + 1] - h_off[a] > h_off[b + 1] - h_off[b]; });
        ANNB_TRY(dmalloc(&ix->d_list_order, nlist, ix));
        ANNB_CUDA_CHECK(cudaMemcpy(ix->d_list_order, order.data(), nlist * 4ull, cudaMemcpyHostToDevice));
    }
    ANNB_TRY(dmalloc(&ix->d_original_ids, std::max<uint64_t>(n_local, 1), ix));
    if (n_local) ANNB_CUDA_CHECK(cudaMemcpyAsync(ix->d_original_ids, original_ids, n_local * 8ull, cudaMemcpyDefault, s));
    if (dtype == ANNB_SQ8) {
        ANNB_TRY(dmalloc(&ix->d_scales, dim, ix));
        ANNB_CUDA_CHECK(cudaMemcpyAsync(ix->d_scales, sq8_scales, dim * 4ull, cudaMemcpyDefault, s));
    }
    ANNB_CUDA_CHECK(cudaStreamSynchronize(s));
    ANNB_TRY(tc_ivf_prepare(ix));
    ANNB_TRY(tc_stream_prepare(ix));
    ANNB_TRY(tc_coarse_prepare(ix));
    ANNB_CUDA_CHECK(cudaStreamSynchronize(s));
    cleanup.armed = false;
    *out = ix;
    return ANNB_OK;
}

int annb_ivf_search(const annb_index* index, const float* queries, uint64_t nq, uint32_t dim, uint32_t k,
                    uint32_t nprobe, uint64_t* out_ids, float* out_dist, uint32_t* out_counts) {
    if (index && dim != index->dim)
        return fail(ANNB_ERR_DIMENSION_MISMATCH, "query dim " + std::to_string(dim) + " != index dim " + std::to_string(index->dim));
    if (index && index->multi) return multi_search(const_cast<annb_index*>(index), true, 0, queries, 0, nq, k, nprobe, out_ids, out_dist, out_counts);
    return search_host(const_cast<annb_index*>(index), true, 0, queries, 0, nq, k, nprobe, 0, out_ids, out_dist, out_counts);
}

int annb_ivf_search_self(const annb_index* index, uint64_t pos_begin, uint64_t pos_end, uint32_t k, uint32_t nprobe,
                         int scatter_to_original, uint64_t* out_ids, float* out_dist, uint32_t* out_counts) {
    if (!index) return fail(ANNB_ERR_INVALID_ARGUMENT, "null index");
    if (index->multi || index->list_begin != 0 || index->list_end != index->nlist) return fail(ANNB_ERR_UNSUPPORTED, "self search needs an unsharded IVF index");
    if (pos_begin > pos_end || pos_end > index->n) return fail(ANNB_ERR_INVALID_ARGUMENT, "position range outside the index");
    return search_host(const_cast<annb_index*>(index), true, 1, nullptr, pos_begin, pos_end - pos_begin, k, nprobe, scatter_to_original,
                       out_ids, out_dist, out_counts);
}

// KnnValidation::validate_index (src/utils/mod.rs:210-242, implemented for the CPU IvfIndex in src/cpu/ivf.rs:496-523): the
// stored vectors at `positions` (internal list-order positions, drawn by the caller -- the reference draws them with its
// StdRng) are searched through the index itself (default nprobe when 0) and exhaustively over the same rows
// (exhaustive_query, :133-196: ties by internal position, results mapped through original_ids); returns the mean of
// |approx ∩ true| / k.  f32 unsharded single-device IVF handles only, like the reference's impl.
int annb_ivf_validate(const annb_index* index, const uint64_t* positions, uint64_t n_samples, uint32_t k, uint32_t nprobe, double* out_recall) {
    annb_index* ix = const_cast<annb_index*>(index);
    if (!ix || !ix->is_ivf) return fail(ANNB_ERR_INVALID_ARGUMENT, "not an IVF index");
    if (!positions || !out_recall || k == 0 || n_samples == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "null buffer / k == 0 / no samples");
    if (ix->multi || ix->list_begin != 0 || ix->list_end != ix->nlist || ix->dtype != ANNB_F32)
        return fail(ANNB_ERR_UNSUPPORTED, "validate_index needs an unsharded single-device f32 IVF index (the reference implements it for IvfIndex<T> only)");
    for (uint64_t i = 0; i < n_samples; i++)
        if (positions[i] >= ix->n) return fail(ANNB_ERR_INVALID_ARGUMENT, "sample position outside the index");
    ANNB_DEVICE(ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    cudaStream_t s = ix->stream;
    ANNB_TRY(order_after_previous(ix, s));
    const uint32_t kk = static_cast<uint32_t>(std::min<uint64_t>(k, ix->n));
    double matches = 0.0;
    DevBuf d_pos, d_q, d_ia, d_it, d_da, d_dt;
    struct Free { DevBuf* b[6]; ~Free() { for (DevBuf* x : b) x->release(); } } guard{{&d_pos, &d_q, &d_ia, &d_it, &d_da, &d_dt}};
    std::vector<uint32_t> pos32;
    std::vector<uint64_t> ia, it;
    for (uint64_t b0 = 0; b0 < n_samples; b0 += QUERY_BATCH) {
        const uint64_t nb = std::min<uint64_t>(QUERY_BATCH, n_samples - b0);
        pos32.resize(nb);
        for (uint64_t i = 0; i < nb; i++) pos32[i] = static_cast<uint32_t>(positions[b0 + i]);
        ANNB_TRY(d_pos.ensure(nb * 4));
        ANNB_TRY(d_q.ensure(nb * ix->row_bytes));
        ANNB_TRY(d_ia.ensure(nb * k * 8ull)); ANNB_TRY(d_it.ensure(nb * k * 8ull));
        ANNB_TRY(d_da.ensure(nb * k * 4ull)); ANNB_TRY(d_dt.ensure(nb * k * 4ull));
        ANNB_CUDA_CHECK(cudaMemcpyAsync(d_pos.p, pos32.data(), nb * 4, cudaMemcpyHostToDevice, s));
        gather_rows_kernel<<<grid_for(nb * (ix->row_bytes >> 4), 256), 256, 0, s>>>(ix->d_rows, ix->row_bytes, d_pos.as<uint32_t>(), static_cast<uint32_t>(nb), d_q.as<uint8_t>());
        ANNB_CUDA_CHECK(cudaGetLastError());
        PreparedQueries pq;                       // f32 rows at the index's own pitch serve as routing and scan queries
        pq.route = d_q.as<float>(); pq.route_ld = ix->row_bytes / 4; pq.scan = d_q.as<uint8_t>(); pq.scan_bytes = ix->row_bytes; pq.qt = QT_F32; pq.bf16_self = 0;
        ANNB_TRY(run_batch(ix, true, pq, nb, k, nprobe, d_ia.as<uint64_t>(), d_da.as<float>(), nullptr, s, true, []() -> int { return ANNB_OK; }));
        ANNB_TRY(flat_simt(ix, pq, nb, k, kk, d_it.as<uint64_t>(), d_dt.as<float>(), nullptr, s, true));   // ids = internal positions
        ia.resize(nb * k); it.resize(nb * k);
        ANNB_CUDA_CHECK(cudaMemcpyAsync(ia.data(), d_ia.p, nb * k * 8ull, cudaMemcpyDeviceToHost, s));
        ANNB_CUDA_CHECK(cudaMemcpyAsync(it.data(), d_it.p, nb * k * 8ull, cudaMemcpyDeviceToHost, s));
        ANNB_CUDA_CHECK(cudaStreamSynchronize(s));
        std::vector<uint64_t> oid(nb * static_cast<uint64_t>(kk));
        {   // original ids of the exhaustive rows (positions -> ids), fetched row by row would be slow: one gather on the device
            DevBuf d_o;
            ANNB_TRY(d_o.ensure(nb * static_cast<uint64_t>(kk) * 8));
            std::vector<uint64_t> flat(nb * static_cast<uint64_t>(kk));
            for (uint64_t q = 0; q < nb; q++) for (uint32_t j = 0; j < kk; j++) flat[q * kk + j] = it[q * k + j];
            cudaError_t e = cudaMemcpyAsync(d_o.p, flat.data(), flat.size() * 8, cudaMemcpyHostToDevice, s);
            if (e == cudaSuccess) { map_ids_kernel<<<grid_for(flat.size(), 256), 256, 0, s>>>(d_o.as<uint64_t>(), flat.size(), ix->d_original_ids, ix->n); e = cudaGetLastError(); }
            if (e == cudaSuccess) e = cudaMemcpyAsync(oid.data(), d_o.p, flat.size() * 8, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            d_o.release();
            ANNB_CUDA_CHECK(e);
        }
        for (uint64_t q = 0; q < nb; q++) {
            uint32_t m = 0;
            for (uint32_t j = 0; j < kk; j++) {
                const uint64_t t = oid[q * kk + j];
                for (uint32_t a = 0; a < k; a++) if (ia[q * k + a] == t) { m++; break; }
            }
            matches += static_cast<double>(m) / static_cast<double>(k);
        }
    }
    *out_recall = matches / static_cast<double>(n_samples);
    return mark_call_done(ix, s);
}

int annb_ivf_search_dev(const annb_index* index, const float* d_queries, uint64_t nq, uint32_t dim, uint32_t k, uint32_t nprobe,
                        uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts, void* stream) {
    annb_index* ix = const_cast<annb_index*>(index);
    if (!ix || !ix->is_ivf) return fail(ANNB_ERR_INVALID_ARGUMENT, "not an IVF index");
    if (dim != ix->dim) return fail(ANNB_ERR_DIMENSION_MISMATCH, "query dim " + std::to_string(dim) + " != index dim " + std::to_string(ix->dim));
    if (!d_queries || !d_out_ids || k == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "null buffer / k == 0");
    if (ix->multi) {
        ANNB_DEVICE(ix->device);
        ANNB_CUDA_CHECK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
        return multi_search(ix, true, 0, d_queries, 0, nq, k, nprobe, d_out_ids, d_out_dist, d_out_counts);
    }
    ANNB_DEVICE(ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ANNB_TRY(order_after_previous(ix, s));
    ix->stat_scanned = 0;
    ix->stat_probed = 0;
    ix->stat_scanned_local = 0;
    ix->stat_tc_tiles = 0;
    for (uint64_t b0 = 0; b0 < nq; b0 += QUERY_BATCH) {
        const uint64_t nb = std::min<uint64_t>(QUERY_BATCH, nq - b0);
        PreparedQueries pq;
        ANNB_TRY(prepare_external(ix, d_queries + b0 * dim, nb, &pq, s));
        ANNB_TRY(run_batch(ix, true, pq, nb, k, nprobe, d_out_ids + b0 * k, d_out_dist ? d_out_dist + b0 * k : nullptr,
                           d_out_counts ? d_out_counts + b0 : nullptr, s, ix->opt_async_dev == 0, []() -> int { return ANNB_OK; }));
    }
    return mark_call_done(ix, s);
}

int annb_ivf_route_dev(const annb_index* index, const float* d_queries, uint64_t nq, uint32_t dim, uint32_t k, uint32_t nprobe,
                       uint32_t* d_probes, uint32_t* d_n_probes, uint32_t probe_pitch, void* stream) {
    annb_index* ix = const_cast<annb_index*>(index);
    if (!ix || !ix->is_ivf || ix->multi) return fail(ANNB_ERR_INVALID_ARGUMENT, "not a (single-device) IVF index");
    if (dim != ix->dim) return fail(ANNB_ERR_DIMENSION_MISMATCH, "query dim " + std::to_string(dim) + " != index dim " + std::to_string(ix->dim));
    if (!d_queries || !d_probes || !d_n_probes || k == 0 || probe_pitch == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "null buffer / k == 0 / zero pitch");
    ANNB_DEVICE(ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ANNB_TRY(order_after_previous(ix, s));
    for (uint64_t b0 = 0; b0 < nq; b0 += QUERY_BATCH) {
        const uint64_t nb = std::min<uint64_t>(QUERY_BATCH, nq - b0);
        PreparedQueries pq;
        ANNB_TRY(prepare_external(ix, d_queries + b0 * dim, nb, &pq, s));
        RouteOut ro{d_probes + b0 * probe_pitch, d_n_probes + b0, probe_pitch};
        ANNB_TRY(route_batch(ix, pq, nb, k, nprobe, ro, s));
    }
    return mark_call_done(ix, s);
}

int annb_ivf_search_probes_dev(const annb_index* index, const float* d_queries, uint64_t nq, uint32_t dim, uint32_t k, uint32_t nprobe,
                               const uint32_t* d_probes, const uint32_t* d_n_probes, uint32_t probe_pitch, uint64_t* d_out_ids, float* d_out_dist,
                               uint32_t* d_out_counts, void* stream) {
    annb_index* ix = const_cast<annb_index*>(index);
    if (!ix || !ix->is_ivf || ix->multi) return fail(ANNB_ERR_INVALID_ARGUMENT, "not a (single-device) IVF index");
    if (dim != ix->dim) return fail(ANNB_ERR_DIMENSION_MISMATCH, "query dim " + std::to_string(dim) + " != index dim " + std::to_string(ix->dim));
    if (!d_queries || !d_probes || !d_n_probes || !d_out_ids || k == 0 || probe_pitch == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "null buffer / k == 0 / zero pitch");
    ANNB_DEVICE(ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ANNB_TRY(order_after_previous(ix, s));
    for (uint64_t b0 = 0; b0 < nq; b0 += QUERY_BATCH) {
        const uint64_t nb = std::min<uint64_t>(QUERY_BATCH, nq - b0);
        PreparedQueries pq;
        ANNB_TRY(prepare_external(ix, d_queries + b0 * dim, nb, &pq, s));
        ANNB_TRY(run_batch(ix, true, pq, nb, k, nprobe, d_out_ids + b0 * k, d_out_dist ? d_out_dist + b0 * k : nullptr,
                           d_out_counts ? d_out_counts + b0 : nullptr, s, ix->opt_async_dev == 0, []() -> int { return ANNB_OK; },
                           d_probes + b0 * probe_pitch, d_n_probes + b0, probe_pitch));
    }
    return mark_call_done(ix, s);
}

// Shard-mode searches (see include/annb200.h): as the _dev searches, but the coverage certificate is left to the caller, who
// knows the merged result: d_out_bound[q] = the distance below which no row of this shard that was NOT re-ranked can lie
// (+inf where every candidate was re-ranked, or the shard ran on the exact kernels).  Nothing is recomputed locally, and
// with preset probes nothing is read back: the call is asynchronous.
static __global__ void fill_f32_kernel(float* __restrict__ p, uint64_t n, float v) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i < n) p[i] = v;
}
static int shard_search(annb_index* ix, bool ivf, const float* d_queries, uint64_t nq, uint32_t dim, uint32_t k, uint32_t nprobe, const uint32_t* d_probes,
                        const uint32_t* d_n_probes, uint32_t probe_pitch, uint64_t* d_out_ids, float* d_out_dist, float* d_out_bound, void* stream) {
    if (!ix || ix->is_ivf != ivf || ix->multi) return fail(ANNB_ERR_INVALID_ARGUMENT, ivf ? "not a (single-device) IVF index" : "not a (single-device) flat index");
    if (dim != ix->dim) return fail(ANNB_ERR_DIMENSION_MISMATCH, "query dim " + std::to_string(dim) + " != index dim " + std::to_string(ix->dim));
    if (!d_queries || !d_out_ids || !d_out_bound || k == 0 || (ivf && (!d_probes || !d_n_probes || probe_pitch == 0))) return fail(ANNB_ERR_INVALID_ARGUMENT, "null buffer / k == 0 / zero pitch");
    ANNB_DEVICE(ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ANNB_TRY(order_after_previous(ix, s));
    fill_f32_kernel<<<grid_for(nq, 256, 1u << 30), 256, 0, s>>>(d_out_bound, nq, INFINITY);   // exact kernels re-rank nothing: their rows are complete
    ANNB_CUDA_CHECK(cudaGetLastError());
    struct Reset { annb_index* ix; ~Reset() { ix->shard_bound = nullptr; } } reset{ix};
    for (uint64_t b0 = 0; b0 < nq; b0 += QUERY_BATCH) {
        const uint64_t nb = std::min<uint64_t>(QUERY_BATCH, nq - b0);
        PreparedQueries pq;
        ANNB_TRY(prepare_external(ix, d_queries + b0 * dim, nb, &pq, s));
        ix->shard_bound = d_out_bound + b0;
        ANNB_TRY(run_batch(ix, ivf, pq, nb, k, nprobe, d_out_ids + b0 * k, d_out_dist ? d_out_dist + b0 * k : nullptr, nullptr, s, false,
                           []() -> int { return ANNB_OK; }, ivf ? d_probes + b0 * probe_pitch : nullptr, ivf ? d_n_probes + b0 : nullptr, probe_pitch));
    }
    return mark_call_done(ix, s);
}

int annb_flat_search_shard_dev(const annb_index* index, const float* d_queries, uint64_t nq, uint32_t dim, uint32_t k, uint64_t* d_out_ids,
                               float* d_out_dist, float* d_out_bound, void* stream) {
    return shard_search(const_cast<annb_index*>(index), false, d_queries, nq, dim, k, 0, nullptr, nullptr, 0, d_out_ids, d_out_dist, d_out_bound, stream);
}

int annb_ivf_search_probes_shard_dev(const annb_index* index, const float* d_queries, uint64_t nq, uint32_t dim, uint32_t k, uint32_t nprobe,
                                     const uint32_t* d_probes, const uint32_t* d_n_probes, uint32_t probe_pitch, uint64_t* d_out_ids, float* d_out_dist,
                                     float* d_out_bound, void* stream) {
    return shard_search(const_cast<annb_index*>(index), true, d_queries, nq, dim, k, nprobe, d_probes, d_n_probes, probe_pitch, d_out_ids, d_out_dist,
                        d_out_bound, stream);
}

// queries whose merged k-th distance does not lie strictly below this shard's bound
static __global__ void shard_check_kernel(const float* __restrict__ bound, const float* __restrict__ merged_dist, uint64_t nq, uint32_t k,
                                          uint32_t* __restrict__ uncert) {
    const uint64_t q = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (q >= nq) return;
    const float dk = merged_dist[q * k + (k - 1)];      // +inf where fewer than k rows were found at all
    if (!(bound[q] > dk) && bound[q] != INFINITY) uncert[1 + atomicAdd(uncert, 1u)] = static_cast<uint32_t>(q);
}

int annb_shard_check_dev(annb_index* ix, const float* d_bound, const float* d_merged_dist, uint64_t nq, uint32_t k, uint32_t* out_count, void* stream) {
    if (!ix || !d_bound || !d_merged_dist || !out_count || k == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "null argument / k == 0");
    if (nq > QUERY_BATCH) return fail(ANNB_ERR_UNSUPPORTED, "shard check: at most 16384 queries per call");
    *out_count = 0;
    ANNB_DEVICE(ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ANNB_TRY(order_after_previous(ix, s));
    ANNB_TRY(ix->s_uncert.ensure((nq + 1) * 4));
    ANNB_CUDA_CHECK(cudaMemsetAsync(ix->s_uncert.p, 0, 4, s));
    shard_check_kernel<<<grid_for(nq, 256, 1u << 30), 256, 0, s>>>(d_bound, d_merged_dist, nq, k, ix->s_uncert.as<uint32_t>());
    ANNB_CUDA_CHECK(cudaGetLastError());
    HostStatus* h = reinterpret_cast<HostStatus*>(ix->h_status);
    ANNB_CUDA_CHECK(cudaMemcpyAsync(&h->n_unc, ix->s_uncert.p, 4, cudaMemcpyDeviceToHost, s));
    ANNB_CUDA_CHECK(cudaStreamSynchronize(s));
    *out_count = h->n_unc;
    ix->stat_uncertified = h->n_unc;
    return mark_call_done(ix, s);
}

// The same test when every shard's bounds travelled with its result block (one process per GPU: the all-gather already
// brought them): each rank evaluates ALL shards' bounds against the merged rows, so every rank derives the same
// "somebody refines" verdict without another collective, and lists its own queries for annb_shard_refine_dev.
static __global__ void shard_check_gathered_kernel(const uint8_t* __restrict__ parts_base, uint64_t stride, uint64_t bound_off, uint32_t parts,
                                                   uint32_t my_part, const float* __restrict__ merged_dist, uint64_t nq, uint32_t k,
                                                   uint32_t* __restrict__ uncert, uint32_t* __restrict__ any) {
    const uint64_t q = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (q >= nq) return;
    if (q == 0) {   // the int32 behind every shard's bounds is its status word: a failed shard poisons the step on every rank alike
        for (uint32_t s = 0; s < parts; s++)
            if (*reinterpret_cast<const int32_t*>(parts_base + s * stride + bound_off + nq * 4) != 0) atomicOr(any, 2u);
    }
    const float dk = merged_dist[q * k + (k - 1)];
    bool some = false;
    for (uint32_t s = 0; s < parts; s++) {
        const float b = reinterpret_cast<const float*>(parts_base + s * stride + bound_off)[q];
        const bool need = !(b > dk) && b != INFINITY;
        some = some || need;
        if (need && s == my_part) uncert[1 + atomicAdd(uncert, 1u)] = static_cast<uint32_t>(q);
    }
    if (some) atomicOr(any, 1u);
}

int annb_shard_check_gathered_dev(annb_index* ix, const void* d_parts, uint64_t part_stride_bytes, uint64_t bound_offset_bytes, uint32_t parts,
                                  uint32_t my_part, const float* d_merged_dist, uint64_t nq, uint32_t k, uint32_t* out_mine, uint32_t* out_any,
                                  void* stream) {
    if (!ix || !d_parts || !d_merged_dist || !out_mine || !out_any || k == 0 || parts == 0 || my_part >= parts) return fail(ANNB_ERR_INVALID_ARGUMENT, "null argument / bad shape");
    if ((part_stride_bytes & 3) || (bound_offset_bytes & 3)) return fail(ANNB_ERR_INVALID_ARGUMENT, "misaligned shard layout");
    if (nq > QUERY_BATCH) return fail(ANNB_ERR_UNSUPPORTED, "shard check: at most 16384 queries per call");
    *out_mine = 0;
    *out_any = 0;
    ANNB_DEVICE(ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ANNB_TRY(order_after_previous(ix, s));
    ANNB_TRY(ix->s_uncert.ensure((nq + 1) * 4));
    ANNB_TRY(ix->s_flags.ensure(64));
    ANNB_CUDA_CHECK(cudaMemsetAsync(ix->s_uncert.p, 0, 4, s));
    ANNB_CUDA_CHECK(cudaMemsetAsync(ix->s_flags.p, 0, 4, s));
    shard_check_gathered_kernel<<<grid_for(nq, 256, 1u << 30), 256, 0, s>>>(static_cast<const uint8_t*>(d_parts), part_stride_bytes, bound_offset_bytes, parts,
                                                                          my_part, d_merged_dist, nq, k, ix->s_uncert.as<uint32_t>(), ix->s_flags.as<uint32_t>());
    ANNB_CUDA_CHECK(cudaGetLastError());
    HostStatus* h = reinterpret_cast<HostStatus*>(ix->h_status);
    ANNB_CUDA_CHECK(cudaMemcpyAsync(&h->n_unc, ix->s_uncert.p, 4, cudaMemcpyDeviceToHost, s));
    ANNB_CUDA_CHECK(cudaMemcpyAsync(&h->overflow, ix->s_flags.p, 4, cudaMemcpyDeviceToHost, s));
    ANNB_CUDA_CHECK(cudaStreamSynchronize(s));
    *out_mine = h->n_unc;
    *out_any = h->overflow;
    ix->stat_uncertified = h->n_unc;
    return mark_call_done(ix, s);
}

// The same check without the read-back: the two verdict words are copied into the caller's (pinned) host buffer on `stream` and
// the call returns at once -- the caller records an event behind it and reads h_verdict[0] (this shard's queries to refine) and
// h_verdict[1] (bit 0: some shard has to refine, bit 1: some shard's call failed) once that event has completed.  Lets a
// serving loop enqueue the next batch before the previous verdict is known (annb200.distributed.ShardedSearch, defer = True).
// The list of queries to refine stays in the handle's scratch only until the handle's next search: a deferred refine repeats
// the (synchronous) check first.
int annb_shard_check_gathered_async_dev(annb_index* ix, const void* d_parts, uint64_t part_stride_bytes, uint64_t bound_offset_bytes, uint32_t parts,
                                        uint32_t my_part, const float* d_merged_dist, uint64_t nq, uint32_t k, uint32_t* h_verdict, void* stream) {
    if (!ix || !d_parts || !d_merged_dist || !h_verdict || k == 0 || parts == 0 || my_part >= parts) return fail(ANNB_ERR_INVALID_ARGUMENT, "null argument / bad shape");
    if ((part_stride_bytes & 3) || (bound_offset_bytes & 3)) return fail(ANNB_ERR_INVALID_ARGUMENT, "misaligned shard layout");
    if (nq > QUERY_BATCH) return fail(ANNB_ERR_UNSUPPORTED, "shard check: at most 16384 queries per call");
    ANNB_DEVICE(ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ANNB_TRY(order_after_previous(ix, s));
    ANNB_TRY(ix->s_uncert.ensure((nq + 1) * 4));
    ANNB_TRY(ix->s_flags.ensure(64));
    ANNB_CUDA_CHECK(cudaMemsetAsync(ix->s_uncert.p, 0, 4, s));
    ANNB_CUDA_CHECK(cudaMemsetAsync(ix->s_flags.p, 0, 4, s));
    shard_check_gathered_kernel<<<grid_for(nq, 256, 1u << 30), 256, 0, s>>>(static_cast<const uint8_t*>(d_parts), part_stride_bytes, bound_offset_bytes, parts,
                                                                          my_part, d_merged_dist, nq, k, ix->s_uncert.as<uint32_t>(), ix->s_flags.as<uint32_t>());
    ANNB_CUDA_CHECK(cudaGetLastError());
    ANNB_CUDA_CHECK(cudaMemcpyAsync(h_verdict, ix->s_uncert.p, 4, cudaMemcpyDeviceToHost, s));
    ANNB_CUDA_CHECK(cudaMemcpyAsync(h_verdict + 1, ix->s_flags.p, 4, cudaMemcpyDeviceToHost, s));
    return mark_call_done(ix, s);
}

int annb_shard_refine_dev(annb_index* ix, const float* d_queries, uint64_t nq, uint32_t dim, uint32_t k, uint32_t nprobe, const uint32_t* d_probes,
                          const uint32_t* d_n_probes, uint32_t probe_pitch, uint64_t* d_ids, float* d_dist, void* stream) {
    if (!ix || ix->multi || !d_queries || !d_ids || k == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "null argument / k == 0");
    if (dim != ix->dim) return fail(ANNB_ERR_DIMENSION_MISMATCH, "query dim " + std::to_string(dim) + " != index dim " + std::to_string(ix->dim));
    if (ix->is_ivf && (!d_probes || !d_n_probes || probe_pitch == 0)) return fail(ANNB_ERR_INVALID_ARGUMENT, "IVF shard: probe lists required");
    if (nq > QUERY_BATCH) return fail(ANNB_ERR_UNSUPPORTED, "shard refine: at most 16384 queries per call");
    ANNB_DEVICE(ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ANNB_TRY(order_after_previous(ix, s));
    HostStatus* h = reinterpret_cast<HostStatus*>(ix->h_status);
    ANNB_CUDA_CHECK(cudaMemcpyAsync(&h->n_unc, ix->s_uncert.p, 4, cudaMemcpyDeviceToHost, s));
    ANNB_CUDA_CHECK(cudaStreamSynchronize(s));
    const uint32_t n_unc = h->n_unc;
    if (n_unc == 0) return ANNB_OK;
    PreparedQueries pq;
    ANNB_TRY(prepare_external(ix, d_queries, nq, &pq, s));
    if (ix->is_ivf) ANNB_TRY(ivf_fallback(ix, pq, k, nprobe, n_unc, d_probes, d_n_probes, probe_pitch, d_ids, d_dist, nullptr, s));
    else ANNB_TRY(flat_fallback(ix, pq, k, n_unc, d_ids, d_dist, nullptr, s));
    return mark_call_done(ix, s);
}

int annb_merge_topk_dev(const uint64_t* d_part_ids, const float* d_part_dist, uint32_t parts, uint64_t nq, uint32_t k,
                        uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts, void* stream) {
    if (!d_part_ids || !d_part_dist || !d_out_ids || parts == 0 || k == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "null buffer / empty shape");
    if (nq == 0) return ANNB_OK;
    MergeParams m{};
    m.part_ids = d_part_ids; m.part_dist = d_part_dist; m.parts = parts; m.k = k; m.nq = nq;
    m.nsort = next_pow2(std::max(parts * k, 2u));
    m.out_ids = d_out_ids; m.out_dist = d_out_dist; m.out_counts = d_out_counts;
    size_t smem = static_cast<size_t>(m.nsort) * 8;
    if (smem > 200 * 1024) return fail(ANNB_ERR_UNSUPPORTED, "merge: parts * k too large");
    ANNB_CUDA_CHECK(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    merge_kernel<<<static_cast<uint32_t>(nq), 128, smem, static_cast<cudaStream_t>(stream)>>>(m);
    ANNB_CUDA_CHECK(cudaGetLastError());
    return ANNB_OK;
}

int annb_merge_shards_dev(const void* d_parts, uint64_t part_stride_bytes, uint64_t dist_offset_bytes, uint32_t parts, uint64_t nq, uint32_t k,
                          uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts, void* stream) {
    if (!d_parts || !d_out_ids || parts == 0 || k == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "null buffer / empty shape");
    if ((part_stride_bytes & 7) || (dist_offset_bytes & 3)) return fail(ANNB_ERR_INVALID_ARGUMENT, "misaligned shard layout");
    if (nq == 0) return ANNB_OK;
    MergeShardsParams m{};
    m.base = static_cast<const uint8_t*>(d_parts); m.part_stride = part_stride_bytes; m.dist_offset = dist_offset_bytes;
    m.parts = parts; m.k = k; m.nq = nq; m.out_ids = d_out_ids; m.out_dist = d_out_dist; m.out_counts = d_out_counts;
    if (m.parts * m.k <= 128u) merge_shards_sort_kernel<false><<<static_cast<uint32_t>(ceil_div<uint64_t>(nq, 4)), 128, 0, static_cast<cudaStream_t>(stream)>>>(m);
    else merge_shards_kernel<<<static_cast<uint32_t>(ceil_div<uint64_t>(nq, 4)), 128, 0, static_cast<cudaStream_t>(stream)>>>(m);
    ANNB_CUDA_CHECK(cudaGetLastError());
    return ANNB_OK;
}

// annb_merge_shards_dev + annb_shard_check_gathered_async_dev in one pass (one kernel, one 8-byte memset, one 8-byte copy to the
// caller's pinned verdict words) for the serving loop's deferred verdict.  The list of this shard's queries to refine is NOT left in
// the handle: a verdict that asks for a refine is followed by the synchronous annb_shard_check_gathered_dev, which builds it.
int annb_merge_check_shards_async_dev(annb_index* ix, const void* d_parts, uint64_t part_stride_bytes, uint64_t dist_offset_bytes, uint64_t bound_offset_bytes,
                                      uint32_t parts, uint32_t my_part, uint64_t nq, uint32_t k, uint64_t* d_out_ids, float* d_out_dist, uint32_t* h_verdict,
                                      void* stream) {
    if (!ix || !d_parts || !d_out_ids || !d_out_dist || !h_verdict || parts == 0 || k == 0 || my_part >= parts) return fail(ANNB_ERR_INVALID_ARGUMENT, "null argument / bad shape");
    if ((part_stride_bytes & 7) || (dist_offset_bytes & 3) || (bound_offset_bytes & 3)) return fail(ANNB_ERR_INVALID_ARGUMENT, "misaligned shard layout");
    if (nq > QUERY_BATCH) return fail(ANNB_ERR_UNSUPPORTED, "shard check: at most 16384 queries per call");
    if (parts * k > 128u || nq == 0) {      // large merges: the two separate passes
        ANNB_TRY(annb_merge_shards_dev(d_parts, part_stride_bytes, dist_offset_bytes, parts, nq, k, d_out_ids, d_out_dist, nullptr, stream));
        return annb_shard_check_gathered_async_dev(ix, d_parts, part_stride_bytes, bound_offset_bytes, parts, my_part, d_out_dist, nq, k, h_verdict, stream);
    }
    ANNB_DEVICE(ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ANNB_TRY(order_after_previous(ix, s));
    ANNB_TRY(ix->s_flags.ensure(64));
    ANNB_CUDA_CHECK(cudaMemsetAsync(ix->s_flags.p, 0, 8, s));
    MergeShardsParams m{};
    m.base = static_cast<const uint8_t*>(d_parts); m.part_stride = part_stride_bytes; m.dist_offset = dist_offset_bytes;
    m.parts = parts; m.k = k; m.nq = nq; m.out_ids = d_out_ids; m.out_dist = d_out_dist; m.out_counts = nullptr;
    m.bound_offset = bound_offset_bytes; m.my_part = my_part; m.verdict = ix->s_flags.as<uint32_t>();
    merge_shards_sort_kernel<true><<<static_cast<uint32_t>(ceil_div<uint64_t>(nq, 4)), 128, 0, s>>>(m);
    ANNB_CUDA_CHECK(cudaGetLastError());
    ANNB_CUDA_CHECK(cudaMemcpyAsync(h_verdict, ix->s_flags.p, 8, cudaMemcpyDeviceToHost, s));
    return mark_call_done(ix, s);
}

int annb_flat_create_multi(annb_index** out, const float* data, uint64_t n, uint32_t dim, int dtype, int metric, const int* devices, int n_devices) {
    if (metric == ANNB_MANHATTAN) return fail(ANNB_ERR_DISTANCE_NOT_SUPPORTED, "Manhattan distance is not supported by the GPU / quantised indices");
    if (metric != ANNB_L2 && metric != ANNB_COSINE) return fail(ANNB_ERR_INVALID_ARGUMENT, "unknown metric");
    if (dtype < ANNB_F32 || dtype > ANNB_SQ8) return fail(ANNB_ERR_INVALID_ARGUMENT, "unknown dtype");
    return multi_flat_create(out, data, n, dim, dtype, metric, devices, n_devices);
}

int annb_ivf_create_multi(annb_index** out, const void* vectors, const void* norms, const float* centroids, const float* centroid_norms,
                          const uint64_t* offsets, const uint64_t* original_ids, uint64_t n, uint32_t dim, uint32_t nlist, int dtype, int metric,
                          const float* sq8_scales, const int* devices, int n_devices) {
    if (metric == ANNB_MANHATTAN) return fail(ANNB_ERR_DISTANCE_NOT_SUPPORTED, "Manhattan distance is not supported by the IVF indices");
    if (metric != ANNB_L2 && metric != ANNB_COSINE) return fail(ANNB_ERR_INVALID_ARGUMENT, "unknown metric");
    return multi_ivf_create(out, vectors, norms, centroids, centroid_norms, offsets, original_ids, n, dim, nlist, dtype, metric, sq8_scales, devices, n_devices);
}

int annb_index_shard_count(const annb_index* ix, uint32_t* out) {
    if (!ix || !out) return fail(ANNB_ERR_INVALID_ARGUMENT, "null argument");
    *out = ix->multi ? static_cast<uint32_t>(ix->multi->shards.size()) : 1u;
    return ANNB_OK;
}

int annb_index_get_info(const annb_index* ix, annb_index_info* out) {
    if (!ix || !out) return fail(ANNB_ERR_INVALID_ARGUMENT, "null argument");
    out->n = ix->n; out->n_total = ix->n_total; out->dim = ix->dim; out->nlist = ix->nlist;
    out->dtype = ix->dtype; out->metric = ix->metric; out->device = ix->device; out->is_ivf = ix->is_ivf;
    out->device_bytes = ix->device_bytes;
    out->host_bytes = sizeof(annb_index) + ix->h_offsets.capacity() * 8;
    return ANNB_OK;
}

int annb_index_set_option(annb_index* ix, const char* key, int64_t value) {
    if (!ix || !key) return fail(ANNB_ERR_INVALID_ARGUMENT, "null argument");
    if (ix->multi) {   // options apply to every shard
        for (annb_index* sh : ix->multi->shards) ANNB_TRY(annb_index_set_option(sh, key, value));
        return ANNB_OK;
    }
    std::lock_guard<std::mutex> lock(ix->mu);
    std::string k(key);
    if (k == "path") { if (value < 0 || value > 2) return fail(ANNB_ERR_INVALID_ARGUMENT, "path must be 0..2"); ix->opt_path = static_cast<int>(value); }
    else if (k == "tc_candidates") ix->opt_tc_candidates = static_cast<int>(value);
    else if (k == "tc_ts") ix->opt_tc_ts = static_cast<int>(value);
    else if (k == "tc_wide_k") ix->opt_tc_wide_k = static_cast<int>(value);
    else if (k == "tc_f32_fp16") {
        // the operand form is chosen when the tensor state is built: rebuild it for a flat f32 index whose setting changes
        const int v = value != 0 ? 1 : 0;
        if (v != ix->opt_tc_f32_fp16) {
            ix->opt_tc_f32_fp16 = v;
            if (!ix->is_ivf && ix->dtype == ANNB_F32 && ix->tc != nullptr) {
                DeviceGuard g(ix->device);
                ANNB_CUDA_CHECK(cudaStreamSynchronize(ix->stream));
                tc_destroy(ix);
                ANNB_TRY(tc_flat_prepare(ix));
                ANNB_CUDA_CHECK(cudaStreamSynchronize(ix->stream));
            } else if (ix->is_ivf && ix->dtype == ANNB_F32 && ix->tc_ivf != nullptr) {
                DeviceGuard g(ix->device);
                ANNB_CUDA_CHECK(cudaStreamSynchronize(ix->stream));
                tc_ivf_destroy(ix);
                ANNB_TRY(tc_ivf_prepare(ix));
                ANNB_CUDA_CHECK(cudaStreamSynchronize(ix->stream));
            }
        }
    }
    else if (k == "tc_strided") ix->opt_tc_strided = static_cast<int>(value);
    else if (k == "tc_f32_lo_smem") ix->opt_tc_f32_lo_smem = static_cast<int>(value);
    else if (k == "tc_bf16_hybrid") ix->opt_tc_bf16_hybrid = static_cast<int>(value);
    else if (k == "tc_bf16_terms") ix->opt_tc_bf16_terms = static_cast<int>(value);
    else if (k == "tc_epi_warps") ix->opt_tc_epi_warps = static_cast<int>(value);
    else if (k == "cert_fallback") ix->opt_cert_fallback = static_cast<int>(value);
    else if (k == "cert_eps_log2") ix->opt_cert_eps = value == 0 ? 0.f : (value > 0 ? -1.0f : std::ldexp(1.0f, static_cast<int>(value)));   // e.g. -18; 0 switches the certificate off; 1 = derived bound (default)
    else if (k == "tc_debug") { DeviceGuard g(ix->device); return ix->is_ivf ? tc_ivf_debug_enable(ix, value != 0) : tc_debug_enable(ix, value != 0); }
    else if (k == "db_splits") ix->opt_db_splits = static_cast<int>(value);
    else if (k == "scan_parts") ix->opt_scan_parts = static_cast<int>(value);
    else if (k == "ivf_task_order") ix->opt_ivf_task_order = static_cast<int>(value);
    else if (k == "ivf_list_major") ix->opt_ivf_list_major = static_cast<int>(value);
    else if (k == "ivf_fast_probe") ix->opt_ivf_fast_probe = static_cast<int>(value);
    else if (k == "ivf_tc_coarse") ix->opt_ivf_tc_coarse = static_cast<int>(value);
    else if (k == "ivf_stream") ix->opt_ivf_stream = static_cast<int>(value);
    else if (k == "ivf_coarse_stage") ix->opt_ivf_coarse_stage = static_cast<int>(value);
    else if (k == "ivf_coarse_gm") ix->opt_ivf_coarse_gm = static_cast<int>(value);
    else if (k == "ivf_task_prefetch") ix->opt_ivf_task_prefetch = static_cast<int>(value);
    else if (k == "ivf_coarse_walk") ix->opt_ivf_coarse_walk = static_cast<int>(value);
    else if (k == "ivf_coarse_blocked") ix->opt_ivf_coarse_blocked = static_cast<int>(value);
    else if (k == "ivf_coarse_fp16") {
        // operand form of the tensor-core centroid ranking (1: 3xFP16, 0: 3xTF32), fixed when its state is built: rebuild on change
        const int v = value != 0 ? 1 : 0;
        if (v != ix->opt_ivf_coarse_fp16) {
            ix->opt_ivf_coarse_fp16 = v;
            if (ix->is_ivf && ix->tc_coarse != nullptr) {
                DeviceGuard g(ix->device);
                ANNB_CUDA_CHECK(cudaStreamSynchronize(ix->stream));
                tc_coarse_destroy(ix);
                ANNB_TRY(tc_coarse_prepare(ix));
                ANNB_CUDA_CHECK(cudaStreamSynchronize(ix->stream));
            }
        }
    }
    else if (k == "async_dev") ix->opt_async_dev = static_cast<int>(value);
    else if (k == "time_kernels") { ix->opt_time_kernels = static_cast<int>(value); ix->timed_ms_total = 0.0; ix->timed_launches = 0; }
    else return fail(ANNB_ERR_INVALID_ARGUMENT, "unknown option " + k);
    return ANNB_OK;
}

int annb_debug_fetch_tile(annb_index* ix, float* host_out) {
    if (!ix || !host_out) return fail(ANNB_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard g(ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    int rc = ix->is_ivf ? tc_ivf_debug_fetch(ix, host_out) : tc_debug_fetch(ix, host_out);
    if (rc != ANNB_OK && rc != ANNB_ERR_CUDA) set_last_error("tc_debug is not enabled on this index");
    return rc;
}

int annb_debug_fetch_uncertified(annb_index* ix, uint32_t* host_out, uint32_t capacity, uint32_t* out_count) {
    if (!ix || !out_count || (!host_out && capacity)) return fail(ANNB_ERR_INVALID_ARGUMENT, "null argument");
    if (ix->multi) return fail(ANNB_ERR_UNSUPPORTED, "per-shard diagnostic");
    *out_count = 0;
    if (!ix->s_uncert.p) return ANNB_OK;
    DeviceGuard g(ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    uint32_t c = 0;
    ANNB_CUDA_CHECK(cudaMemcpy(&c, ix->s_uncert.p, 4, cudaMemcpyDeviceToHost));
    *out_count = c;
    const uint32_t m = std::min(c, capacity);
    if (m) ANNB_CUDA_CHECK(cudaMemcpy(host_out, ix->s_uncert.as<uint32_t>() + 1, m * 4ull, cudaMemcpyDeviceToHost));
    return ANNB_OK;
}

int annb_debug_fetch_cycles(annb_index* ix, uint64_t* host_out8) {
    if (!ix || !host_out8) return fail(ANNB_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard g(ix->device);
    std::lock_guard<std::mutex> lock(ix->mu);
    int rc = ix->is_ivf ? tc_ivf_debug_cycles(ix, reinterpret_cast<unsigned long long*>(host_out8)) : tc_debug_cycles(ix, reinterpret_cast<unsigned long long*>(host_out8));
    if (rc != ANNB_OK && rc != ANNB_ERR_CUDA) set_last_error("tc_debug is not enabled on this index");
    return rc;
}

int annb_index_get_stat(const annb_index* ix, const char* key, int64_t* out) {
    if (!ix || !key || !out) return fail(ANNB_ERR_INVALID_ARGUMENT, "null argument");
    if (ix->multi) {   // counters add up over the shards; path-like values are the first shard's
        const std::string kk(key);
        const bool sum = kk != "last_path" && kk != "coarse_path";
        int64_t acc = 0;
        for (size_t i = 0; i < ix->multi->shards.size(); i++) {
            int64_t v = 0;
            ANNB_TRY(annb_index_get_stat(ix->multi->shards[i], key, &v));
            if (sum) acc += v; else if (i == 0) acc = v;
        }
        *out = acc;
        return ANNB_OK;
    }
    std::string k(key);
    if (k == "kernel_launches") *out = ix->stat_launches;
    else if (k == "scanned_vectors") *out = ix->stat_scanned;
    else if (k == "probed_lists") *out = ix->stat_probed;
    else if (k == "scanned_vectors_local") *out = ix->stat_scanned_local;
    else if (k == "tc_scan_tiles") *out = ix->stat_tc_tiles;
    else if (k == "last_path") *out = ix->stat_last_path;
    else if (k == "coarse_path") *out = ix->stat_coarse_path;
    else if (k == "fallback_queries") *out = ix->stat_fallback_queries;
    else if (k == "cert_eps_bits") *out = ix->stat_cert_eps_bits;
    else if (k == "tc_escalated") *out = ix->tc_escalate;
    else if (k == "tc_kind") *out = ix->is_ivf ? tc_ivf_kind(ix) : tc_flat_kind(ix);
    else if (k == "uncertified") {
        // queries of the last tensor-path call whose pre-selection margin could not be certified (see rerank_kernel)
        *out = 0;
        if (ix->s_uncert.p && ix->stat_last_path == ANNB_PATH_TENSOR) {
            DeviceGuard g(ix->device);
            std::lock_guard<std::mutex> lock(ix->mu);
            uint32_t c = 0;
            ANNB_CUDA_CHECK(cudaMemcpy(&c, ix->s_uncert.p, 4, cudaMemcpyDeviceToHost));
            *out = c;
        }
    }
    else if (k == "dominant_kernel_ns" || k == "dominant_kernel_launches") {
        // synchronises on the recorded events; total device time of the dominant kernel since time_kernels was set
        DeviceGuard g(ix->device);
        std::lock_guard<std::mutex> lock(ix->mu);
        collect_timers(ix);
        *out = (k == "dominant_kernel_ns") ? static_cast<int64_t>(ix->timed_ms_total * 1e6) : ix->timed_launches;
    }
    else return fail(ANNB_ERR_INVALID_ARGUMENT, "unknown stat " + k);
    return ANNB_OK;
}

}  // extern "C"
