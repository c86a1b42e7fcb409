// flat_tc.hpp -- host interface of the tensor-core (tcgen05 / TMEM / TMA) flat search path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct annb_index;

namespace annb {

// Builds the tensor-core operand copies (tf32 hi/lo split, norms) and TMA descriptors of a flat index.
int tc_flat_prepare(annb_index* ix);
// True if (dtype, dim, k, query type) is covered by the tensor path.
bool tc_flat_supported(const annb_index* ix, int qt, uint32_t k_eff);
// Flat search on the tensor path: approximate pre-selection of k' candidates per (query, split) on the
// tensor cores, then exact re-rank in the reference's arithmetic and merge.
int tc_flat_search(annb_index* ix, const uint8_t* d_q, uint32_t q_bytes, int qt, int bf16_self, uint64_t nq, uint32_t k_eff,
                   uint32_t k_out, uint64_t* d_ids, float* d_dist, uint32_t* d_cnt, cudaStream_t s);
void tc_destroy(annb_index* ix);
int tc_flat_kind(const annb_index* ix);   // operand form: -1 none, 0 3xTF32, 1 bf16 terms, 2 int8, 3 3xFP16
// Error bound the coverage certificates assume for the pre-selection values of a (kind, padded K, query terms) kernel.
uint32_t tc_bf16_terms(const annb_index* ix);
float tc_cert_eps(const annb_index* ix, int kind, uint32_t kp_elems, uint32_t terms, bool inkernel_split);

// Tensor-core IVF list scan (ivf_tc.cu): operand copies in list order, grouped (list x 128 queries) tasks, exact re-rank.
int tc_ivf_prepare(annb_index* ix);
void tc_ivf_destroy(annb_index* ix);
bool tc_ivf_supported(const annb_index* ix, int qt, uint32_t k_eff);
int tc_ivf_kind(const annb_index* ix);    // operand form of the tensor scan: -1 none, 0 3xTF32, 1 bf16 terms, 2 int8, 3 3xFP16
uint32_t tc_ivf_kprime(const annb_index* ix, uint32_t k_eff);   // candidates kept per (query, rank, half tile)
int tc_ivf_scan(annb_index* ix, const uint8_t* d_q, uint32_t q_bytes, uint64_t nq, uint32_t k_eff, uint32_t k_out, uint32_t probe_pitch,
                const uint32_t* d_pair_off, const uint32_t* d_task_off, const void* d_pairs, uint32_t* d_task_counter, uint64_t max_tasks,
                const uint32_t* d_n_probes, const uint64_t* row_map, uint64_t* d_ids, float* d_dist, uint32_t* d_cnt, cudaStream_t s,
                const void* d_tasks);
// HBM-streaming query-major list scan (ivf_stream.cu): per-warp rings of TMA tile loads over the index rows themselves.
struct StreamState;
struct StreamScanArgs {
    const uint8_t* queries; uint32_t q_bytes; uint64_t nq; int qt; int bf16_self;
    const uint32_t* probes; uint32_t probe_pitch; const uint32_t* n_probes;
    uint32_t parts, subs, k, nsort;
    uint64_t* part_keys;
};
int tc_stream_prepare(annb_index* ix);
void tc_stream_destroy(annb_index* ix);
bool tc_stream_supported(const annb_index* ix);
int tc_stream_scan(annb_index* ix, const StreamScanArgs& a, cudaStream_t s);
// Tensor-core centroid ranking of an IVF index: dense approximate values (DENSE mode of the flat kernel) + per-query
// radix select, exact re-computation and certification of the `pitch` nearest cells.
int tc_coarse_prepare(annb_index* ix);
void tc_coarse_destroy(annb_index* ix);
bool tc_coarse_supported(const annb_index* ix);
// Probe expansion fused into the select (select_probed_clusters, src/utils/k_means_utils.rs:3007-3029; the rule of probe_walk_kernel):
// with `walk` the select's warp walks its own certified prefix and writes the probe list -- d_ranked is then not written at all.
struct CoarseWalk {
    const uint64_t* offsets;           // [nlist + 1] global list offsets
    uint32_t nprobe;
    uint64_t k;
    uint32_t* probes;                  // [nq][pitch]
    uint32_t* n_probes;                // [nq]
    uint32_t* overflow;                // set to 1 when a query's prefix ends before the rule is satisfied
    unsigned long long* stat_scanned;  // sum over queries of reachable vectors
    unsigned long long* stat_probed;   // [0] probed cells, [1] reachable vectors of this shard's lists
    uint32_t list_begin, list_end;
};
// Uniform power-of-two scale of a 3xFP16 L2 database operand: the largest finite |element| of the `count` floats lands in [2^13, 2^14).
int tc_uniform_f16_scale(annb_index* ix, const float* d_values, uint64_t count, float* out_scale);
int tc_coarse_rank(annb_index* ix, const float* d_route, uint32_t route_ld, uint64_t nq, uint32_t pitch, uint64_t* d_ranked, cudaStream_t s,
                   const CoarseWalk* walk = nullptr);
// Build-side assignment (assign_all_parallel) on the tensor path: the centroid table searched like a flat f32 index with
// k' = 16, exact scores of the survivors in the reference's arithmetic, certificate; uncertified rows are listed for the
// exact CUDA-core kernel.  tc_assign_create leaves *out == nullptr for shapes it does not cover.
struct TcAssignState;
int tc_assign_create(TcAssignState** out, uint32_t dim, uint32_t nlist);
void tc_assign_destroy(TcAssignState* st);
int tc_assign_set_centroids(TcAssignState* st, const float* d_c, uint32_t c_ld, const float* d_assign_aux, bool cosine, cudaStream_t s);
int tc_assign_run(TcAssignState* st, const float* d_x, uint32_t x_ld, uint64_t nr, const float* d_c, uint32_t c_ld, const float* d_assign_aux,
                  bool cosine, uint32_t* d_assign, uint32_t* d_uncert, cudaStream_t s);
// Test hooks: CTA (0,0) of the tensor kernel dumps the 128 x 128 values of its first tile.
int tc_debug_enable(annb_index* ix, bool on);
int tc_debug_fetch(annb_index* ix, float* host_out);
int tc_ivf_debug_fetch(annb_index* ix, float* host_out);
int tc_ivf_debug_enable(annb_index* ix, bool on);
int tc_ivf_debug_cycles(annb_index* ix, unsigned long long* host_out8);
int tc_debug_cycles(annb_index* ix, unsigned long long* host_out8);

}  // namespace annb
