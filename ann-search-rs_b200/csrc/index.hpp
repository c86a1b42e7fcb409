// index.hpp -- host-side state behind the opaque annb_index handle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

struct annb_multi;

namespace annb {

// Grow-only device scratch buffer.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return ANNB_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + (bytes >> 3) + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            set_last_error(std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e));
            (void)cudaGetLastError();
            return e == cudaErrorMemoryAllocation ? ANNB_ERR_OUT_OF_MEMORY : ANNB_ERR_CUDA;
        }
        cap = want;
        return ANNB_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct TcState;     // tensor-core operand copies + tensor maps (flat_tc.cu)
struct StreamState; // tensor map of the streaming list scan (ivf_stream.cu)
struct IvfTcState;  // same for the IVF list scan (ivf_tc.cu)

}  // namespace annb

struct annb_index {
    int device = 0;
    int dtype = ANNB_F32;
    int metric = ANNB_L2;
    bool is_ivf = false;
    uint64_t n = 0;        // rows stored by this handle
    uint64_t n_total = 0;  // rows of the whole index
    uint32_t dim = 0;
    uint32_t row_bytes = 0;  // padded row pitch of d_rows
    uint64_t id_base = 0;

    uint8_t* d_rows = nullptr;     // [n][row_bytes] in the index dtype
    float* d_norms = nullptr;      // [n] cosine (f32 / bf16)
    int32_t* d_norms_i = nullptr;  // [n] cosine (sq8)
    float* d_scales = nullptr;     // [dim] sq8

    // IVF
    uint32_t nlist = 0, list_begin = 0, list_end = 0;
    uint32_t cent_ld = 0;             // centroid pitch in floats
    float* d_centroids = nullptr;     // [nlist][cent_ld]
    float* d_centroid_norms = nullptr;
    uint64_t* d_offsets = nullptr;       // [nlist+1] global
    uint64_t* d_original_ids = nullptr;  // [n]
    uint32_t* d_list_order = nullptr;    // [nlist] lists by descending length (task order of the tensor-core scan)
    uint64_t shard_row0 = 0;
    std::vector<uint64_t> h_offsets;

    // options
    int opt_path = ANNB_PATH_AUTO;
    int opt_tc_candidates = 0;
    int opt_tc_bf16_hybrid = 0; // flat tensor path, bf16 index + f32 queries: third query term in shared memory (SS-mode MMA) instead of TMEM
    int opt_tc_bf16_terms = 0;  // tensor paths, bf16 index + f32 queries: bf16 terms the query is split into (2 or 3); 0 = by metric (cosine 2, L2 3)
    int opt_tc_epi_warps = 0;   // flat tensor path, bf16 / int8 kernels: epilogue warps per TMEM lane quarter (0 = auto: four for k' = 16, 2 = two)
    int tc_escalate = 0;        // sticky level: a batch of this flat handle left more than 2 % of its queries uncertified -> 1: k' = 32, 2: wide-k mode
    int opt_tc_wide_k = 1;      // flat tensor path: serve 24 < k <= 128 from the union of interleaved k' = 32 lists (0: such k go to the CUDA-core path)
    int opt_tc_strided = 0;     // flat tensor path: interleave the splits' tiles over the database also for k <= 24
    int opt_tc_f32_lo_smem = 0; // flat tensor path, f32 rows of <= 128 elements: lo query piece in shared memory (frees TMEM for a third accumulator stage)
    int opt_tc_f32_fp16 = 1;    // flat tensor path, f32 index, rows <= 256 elements: 3xFP16 (rows scaled by powers of two) instead of 3xTF32
    int opt_tc_ts = 1;         // tensor path, f32: keep the query operand in TMEM (TS-mode MMA)
    int opt_db_splits = 0;
    int opt_scan_parts = 0;
    int opt_ivf_fast_probe = 1;   // rank only nprobe + 64 centroids with the fused select (0: always dense matrix + full sort)
    int opt_ivf_task_order = 1;   // tensor-core IVF scan: 1 = tasks handed out longest list first, 0 = in list order
    int opt_ivf_list_major = -1;  // -1 auto, 0 query-major scan, 1 list-major scan
    int opt_time_kernels = 0;  // record CUDA events around the dominant kernel of every search

    // stats
    mutable int64_t stat_launches = 0;
    mutable int64_t stat_scanned = 0, stat_probed = 0, stat_last_path = 0, stat_uncertified = 0;
    mutable int64_t stat_scanned_local = 0;   // vectors of this shard's own lists among stat_scanned
    mutable int64_t stat_tc_tiles = 0;      // last host-buffer call: 128-row tiles executed by the tensor-core IVF scan (padded MMA work = tiles * 128 * 128 * kp)

    uint64_t device_bytes = 0;

    // dominant-kernel timing (option "time_kernels"): event pairs recorded on the launching stream
    mutable std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;
    mutable double timed_ms_total = 0.0;
    mutable int64_t timed_launches = 0;

    // scratch (guarded by mu)
    mutable std::mutex mu;
    cudaStream_t stream = nullptr;
    annb::DevBuf s_qpad, s_qcodes, s_route, s_cdist, s_probes, s_nprobes, s_keys, s_flags, s_ids, s_dist, s_cnt, s_tmp, s_pairs;
    float tc_xnorm_max = 0.f;      // largest stored-row norm (coverage certificate of the tensor paths, L2)
    float opt_cert_eps = -1.0f;    // < 0: derived per kernel from its MMA count (tc_cert_eps); >= 0: caller override, 0 = certificate off
    int opt_cert_fallback = 1;     // re-run uncertified queries on the exact CUDA-core path
    mutable int64_t stat_fallback_queries = 0;  // cumulative
    mutable int64_t stat_cert_eps_bits = 0;     // f32 bit pattern of the error bound the last tensor-path certificate used
    float* shard_bound = nullptr;  // set (under mu) for the duration of a shard-mode search: the re-rank kernel reports per-query bounds there
    void* h_status = nullptr;      // pinned host block the status words of a batch are read back into (api.cu HostStatus)
    cudaEvent_t last_event = nullptr;   // end of the last call that used the scratch buffers, and the stream it ran on
    bool last_event_valid = false;
    cudaStream_t last_stream = nullptr;
    int opt_async_dev = 0;         // 1: the _dev entry points never synchronise (no certificate fallback, no ranking retry; the caller polls the stats)
    struct annb_multi* multi = nullptr;   // multi-device handle (multi.cu): this object is only the front of its per-device shards
    annb::DevBuf s_fbq, s_fbr, s_fbi, s_fbd, s_fbc;
    annb::DevBuf s_uncert;         // [1 + nq] uncertified-query counter + list of the last tensor-path call
    annb::TcState* tc = nullptr;
    annb::IvfTcState* tc_ivf = nullptr;
    annb::StreamState* tc_stream = nullptr;
    int opt_ivf_coarse_stage = 0;         // tensor-core centroid ranking: stage each query's value row in shared memory (measured slower, off)
    int opt_ivf_task_prefetch = 1;        // IVF tensor scan: next task claimed / fetched by the producer lane a few tiles before the current one ends
    int opt_ivf_coarse_gm = 1;            // tensor-core centroid ranking: select from the dense kernel's group minima (coarse_select_gm_kernel)
    int opt_ivf_coarse_walk = 1;          // tensor-core centroid ranking: probe expansion inside the select kernel (0: separate probe_walk_kernel)
    int opt_ivf_coarse_blocked = 1;       // ... with the dense matrix in the blocked layout (coalesced epilogue stores); 0: row-major
    int opt_ivf_coarse_fp16 = 1;          // tensor-core centroid ranking: 3xFP16 operands (0: 3xTF32); read once, when the index is created
    int opt_ivf_stream = 1;               // query-major list scan: 1 = TMA-ring streaming kernel (ivf_stream.cu), 0 = cp.async kernel
    annb::TcState* tc_coarse = nullptr;   // centroid table as a tensor-core operand (IVF centroid ranking)
    float tc_cnorm_max = 0.f;             // largest centroid norm
    int opt_ivf_tc_coarse = 1;            // rank the centroids on the tensor cores (0: CUDA-core ranking only)
    mutable int64_t stat_coarse_path = 0; // last IVF call: 0 exact dense ranking, 1 fused CUDA-core select, 2 tensor cores
};
