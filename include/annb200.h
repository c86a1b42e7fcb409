/*
 * annb200.h -- C ABI of libannb200: B200 (sm_100a) flat and IVF kNN search.
 *
 * This is the drop-in boundary for the data-parallel search hot path of
 * GregorLueg/ann-search-rs.  The reference has no FFI of its own: the seam is
 * the `*_gpu` family of free functions in src/lib.rs:2813-3002 whose device is a
 * CubeCL `R::Device` type parameter, plus the BF16 / SQ8 twins in
 * src/lib.rs:1702-1871 and 2100-2290.  Each entry point below names the
 * reference item it replaces (paths relative to the reference crate root).
 * INTEGRATION.md shows the Rust `extern "C"` block and the safe wrappers that
 * keep the crate's build_and query_ signatures on top of these symbols.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross the boundary;
 *   - every function returns an annb_status (0 = OK, negative = error class that
 *     maps 1:1 onto a variant of AnnSearchErrors, src/errors.rs); the message of
 *     the last failure on the calling thread is available via annb_last_error();
 *   - host-buffer entry points (`annb_*_search`) accept pageable or pinned host
 *     memory, or device memory (resolved through UVA);
 *   - `_dev` entry points take device pointers and a cudaStream_t (as void*): all work is enqueued on that stream and the
 *     results are complete, in stream order, when the call returns.  The call itself may synchronise the stream once
 *     per internal batch of 16384 queries: the tensor-core paths are optimistic, and a 40-byte status block (queries that
 *     failed the coverage certificate, a probe set that outgrew its ranked prefix) is read back to decide whether the
 *     exact kernels have to run for some queries.  Option "async_dev" = 1 drops that read-back: the call then never
 *     blocks, nothing is recomputed, and the caller polls get_stat "uncertified" after its own synchronisation;
 *   - calls on one handle may come from several threads and on different streams: a per-handle mutex orders the host
 *     side, and a call whose stream differs from the previous call's first waits (on the device) for that call's end,
 *     because the scratch buffers are shared;
 *   - the create / host-buffer entry points copy on a private non-blocking
 *     stream: device buffers passed to them must be complete (no kernel of
 *     the caller still writing them) when the call is made;
 *   - results: `out_ids` is [nq * k] uint64, `out_dist` is [nq * k] float
 *     (may be NULL = return_dist false), `out_counts` [nq] uint32 (may be NULL).
 *     Row i holds counts[i] = min(k, reachable) valid entries in ascending
 *     (distance, id-within-index) order; the tail is padded with
 *     id = UINT64_MAX, dist = +inf.  Distances are squared Euclidean, or
 *     1 - cos; SQ8 distances are in code space (src/utils/dist.rs:5015-5077);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     returns ANNB_ERR_CUDA.
 */
#ifndef ANNB200_H
#define ANNB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANNB_VERSION_MAJOR 0
#define ANNB_VERSION_MINOR 1

typedef struct annb_index annb_index; /* opaque; device memory is owned by the library */

/* Storage type of the database vectors (queries are always f32 at the ABI). */
enum annb_dtype {
    ANNB_F32 = 0,  /* ExhaustiveIndex / IvfIndex              (src/cpu/exhaustive.rs, src/cpu/ivf.rs)        */
    ANNB_BF16 = 1, /* ExhaustiveIndexBf16 / IvfIndexBf16      (src/quantised/{exhaustive,ivf}_bf16.rs)       */
    ANNB_SQ8 = 2   /* ExhaustiveSq8Index / IvfSq8Index        (src/quantised/{exhaustive,ivf}_sq8.rs)        */
};

/* Dist::{SquaredEuclidean, Cosine} (src/utils/dist.rs:29-37).  Manhattan is rejected
 * by every GPU / quantised / IVF constructor of the reference
 * (src/gpu/exhaustive_gpu.rs:73-75, src/cpu/ivf.rs:153-155) and here. */
enum annb_metric { ANNB_L2 = 0, ANNB_COSINE = 1, ANNB_MANHATTAN = 2 };

enum annb_status {
    ANNB_OK = 0,
    ANNB_ERR_DIMENSION_MISMATCH = -1,     /* AnnSearchErrors::DimensionMismatch         src/errors.rs:23  */
    ANNB_ERR_DISTANCE_NOT_SUPPORTED = -2, /* AnnSearchErrors::DistanceNotSupported      src/errors.rs:32  */
    ANNB_ERR_TOO_FEW_SAMPLES = -3,        /* AnnSearchErrors::TooFewSamplesForCentroids src/errors.rs:90  */
    ANNB_ERR_INVALID_ARGUMENT = -4,       /* null pointer, n == 0, unsupported k / nlist                 */
    ANNB_ERR_CUDA = -5,                   /* replaces CubeClServerError / CubeclUtils   src/errors.rs:212-223 */
    ANNB_ERR_NCCL = -6,
    ANNB_ERR_OUT_OF_MEMORY = -7,
    ANNB_ERR_UNSUPPORTED = -8             /* e.g. DimTooHighForSharedMemory             src/errors.rs:231 */
};

/* Search-path selector (annb_index_set_option "path"). */
enum annb_path { ANNB_PATH_AUTO = 0, ANNB_PATH_SIMT = 1, ANNB_PATH_TENSOR = 2 };

const char* annb_last_error(void);
int annb_version(void); /* major * 1000 + minor */
int annb_device_count(int* out);

/* parse_ann_dist (src/utils/dist.rs:63-70): "euclidean"|"l2", "cosine", "manhattan"|"l1",
 * case-insensitive.  Returns the annb_metric, or -1 for an unknown string (the reference's
 * free functions then warn and fall back to L2, src/lib.rs:274-277 -- that policy stays on
 * the caller's side). */
int annb_parse_metric(const char* s);

/* ---------------------------------------------------------------- ingest -- */

/* matrix_to_flat (src/utils/mod.rs:44-68) on the device: a strided samples x features view (faer MatRef: element (r, c) at
 * mat[r * row_stride + c * col_stride], strides in elements; faer's own matrices are column-major, row_stride = 1) ->
 * contiguous row-major f32 [nrows * ncols].  `mat` and `out_rowmajor` may each be host or device memory; a device result
 * can be handed straight to annb_flat_create / annb_ivf_assign / annb_kmeans_lloyd, which accept device pointers for
 * their row-major inputs.  The reference does this with a single-threaded strided gather on the host; here a
 * column-major host matrix is packed by a strided copy on the way up and transposed in 32 x 32 shared-memory tiles.
 * Negative strides (reversed views) are not supported. */
int annb_matrix_to_flat(const float* mat, uint64_t nrows, uint32_t ncols, int64_t row_stride, int64_t col_stride,
                        float* out_rowmajor, int device);

/* Host-only helpers for the crate's on-disk format (src/serialise/mod.rs: bincode 2 "standard" configuration): the
 * variable-length encoding of its Vec<usize> fields.  encode returns the bytes written, decode the bytes consumed for
 * `count` values; -1 on a short buffer or a marker byte that 64-bit values never produce.  No device is touched. */
int64_t annb_varint_encode_u64(const uint64_t* values, uint64_t count, uint8_t* out, uint64_t out_capacity);
int64_t annb_varint_decode_u64(const uint8_t* buf, uint64_t len, uint64_t count, uint64_t* out);

/* ------------------------------------------------------------------ flat -- */

/* Replaces ExhaustiveIndexGpu::new (src/gpu/exhaustive_gpu.rs:72-110), and for dtype BF16 / SQ8
 * ExhaustiveIndexBf16::new (src/quantised/exhaustive_bf16.rs:94-121) and ExhaustiveSq8Index::new
 * (src/quantised/exhaustive_sq8.rs:104-151).  `data` is the row-major f32 matrix that
 * matrix_to_flat (src/utils/mod.rs:44-68) produces; norms, BF16 rounding and SQ8
 * normalise/train/encode run on the device with the reference's arithmetic.
 * `sq8_scales` (dim floats) may be given to reuse a codebook (shards of one index); NULL trains
 * on `data`.  `id_base` is added to every returned id (row-range shards).  The database stays
 * resident on `device` for the life of the handle (the reference re-uploads it per call,
 * src/gpu/dist_gpu.rs:662). */
int annb_flat_create(annb_index** out, const float* data, uint64_t n, uint32_t dim, int dtype, int metric,
                     const float* sq8_scales, uint64_t id_base, int device);

/* Replaces ExhaustiveIndexGpu::query_batch (src/gpu/exhaustive_gpu.rs:124-162) /
 * query_exhaustive_index_gpu (src/lib.rs:2842) and the BF16 / SQ8 query paths
 * (src/lib.rs:1733, 1818).  queries: [nq * dim] f32 row-major. */
int annb_flat_search(const annb_index* index, const float* queries, uint64_t nq, uint32_t dim, uint32_t k,
                     uint64_t* out_ids, float* out_dist, uint32_t* out_counts);

/* Replaces ExhaustiveIndexGpu::generate_knn (src/gpu/exhaustive_gpu.rs:179-198) and the
 * generate_knn of the CPU / BF16 / SQ8 flat indices: every stored row queries the index
 * (self included at rank 0).  Outputs are [n * k].  With row_begin/row_end a sub-range of rows
 * acts as the query set (outputs [(row_end-row_begin) * k]); pass 0, n for the whole index. */
int annb_flat_search_self(const annb_index* index, uint64_t row_begin, uint64_t row_end, uint32_t k,
                          uint64_t* out_ids, float* out_dist, uint32_t* out_counts);

/* Exact kNN graph rows in the shape of KnnGraphGpu (src/gpu/nndescent_gpu.rs:2418-2446), the hand-off struct
 * build_nsg_from_gpu_knn (src/lib.rs:3330-3345 -> NsgIndex::build_from_knn, src/cpu/nsg.rs:744-775) and the raw-kNN
 * consumers read: for every stored row in [row_begin, row_end) its k nearest OTHER rows -- the self edge is dropped by
 * id, as compact_knn_rows does (nndescent_gpu.rs:2631-2667) -- ascending by distance, unfilled slots padded with the
 * reference's sentinel pair (SENTINEL_PID = u32::MAX >> 1, f32::MAX).  out_pid / out_dist are [(row_end - row_begin) * k];
 * out_counts (may be NULL) receives the number of real neighbours per row.  The reference fills this struct with
 * NN-Descent (approximate); here it is the exhaustive self search of annb_flat_search_self with k + 1, so the graph is
 * exact.  f32 indices only (the struct carries the f32 vectors); single- and multi-device handles. */
int annb_flat_knn_graph(const annb_index* index, uint64_t row_begin, uint64_t row_end, uint32_t k, uint64_t* out_pid,
                        float* out_dist, uint32_t* out_counts);

/* Device-pointer variants (inputs already resident in HBM; asynchronous on `stream`). */
int annb_flat_search_dev(const annb_index* index, const float* d_queries, uint64_t nq, uint32_t dim, uint32_t k,
                         uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts, void* stream);

/* ------------------------------------------------------------------- IVF -- */

/* Coarse assignment used by the IVF constructors: replaces assign_all_gpu
 * (src/gpu/k_means_gpu.rs:2658-2726) / assign_all_parallel (src/utils/k_means_utils.rs:2214-2241):
 * argmax_c 2 x.c - |c|^2 (L2) or x.c / |c| (cosine), lowest centroid id on ties.
 * data [n*dim], centroids [nlist*dim] (host or device), out_assign [n] uint32 (host).
 * centroid_norms [nlist] is what the caller of direct_assign passes (cosine only): the
 * sequential-fold norms of IvfIndex::build (src/cpu/ivf.rs:193-206) or all ones for the SQ8
 * index (src/quantised/ivf_sq8.rs:214-215); NULL computes the former on the device. */
int annb_ivf_assign(const float* data, uint64_t n, uint32_t dim, const float* centroids,
                    const float* centroid_norms, uint32_t nlist, int metric, uint32_t* out_assign, int device);
/* Tables of >= 512 centroids with rows of <= 128 f32 elements (and n >= 4096) are assigned on the tensor cores: the centroid
 * table is searched like a flat f32 index (3xTF32 values of |c|^2 - 2 x.c or -x.c/|c|, 16 cells kept per row), the kept
 * cells' scores are recomputed in the direct_assign arithmetic, and a row whose winner cannot be certified against the
 * pruning threshold is redone on the exact CUDA-core kernel -- the assignments are the same bits either way.
 * annb_assign_last_redone: rows of this thread's last annb_ivf_assign / annb_kmeans_lloyd call that were redone
 * (diagnostic).  The environment variable ANNB200_ASSIGN_PATH=simt keeps both calls on the exact kernel. */
uint64_t annb_assign_last_redone(void);

/* Lloyd iterations of train_centroids on the device (first step of section 8f: IVF build on the GPU).  Restates the
 * unbalanced `parallel_lloyd` loop (src/utils/k_means_utils.rs:1572-1700; GPU analogue src/gpu/k_means_gpu.rs:1813-2260):
 * per iteration  assignment with the direct_assign arithmetic (bit-identical to annb_ivf_assign; cosine uses
 * calculate_l2_norm of the current centroids)  ->  stop if at most max(1, n / 10000) assignments changed (tested before
 * the update)  ->  centroid = mean of its members (f64 accumulation; empty clusters keep their centroid).
 * `centroids` [nlist * dim] (host or device): in = the initial centroids (the reference draws them with rand's StdRng,
 * fast_random_init / k-means||, which stays on the caller's side), out = the trained centroids.  `out_iters` (may be
 * NULL): number of centroid updates performed.  data [n * dim] host or device. */
int annb_kmeans_lloyd(const float* data, uint64_t n, uint32_t dim, float* centroids, uint32_t nlist, int metric,
                      uint32_t max_iters, uint32_t* out_iters, int device);

/* The same loop with the balancing hook of KMeansTrainingParams::with_balancing (src/utils/k_means_utils.rs:286-345):
 * after every mean update adjust_centers (:979-1030; RAFT's balanced k-means) pulls each centroid whose cluster holds at
 * most a quarter of the average size toward a point of an above-average cluster (weight min(count, 5); empty clusters
 * jump onto the donor), donors found by the reference's strided walk seeded with seed + iteration; the loop only stops
 * once an iteration moved no centroid (:1618).  No random numbers are involved, so the result is reproducible and is
 * tested against the oracle's restatement.  out_adjusted (may be NULL): total number of centroid moves. */
int annb_kmeans_lloyd_balanced(const float* data, uint64_t n, uint32_t dim, float* centroids, uint32_t nlist, int metric,
                               uint32_t max_iters, int balanced, uint64_t seed, uint32_t* out_iters, uint64_t* out_adjusted,
                               int device);

/* Builds a resident IVF index from the contents of the reference's index struct after
 * optimise_memory_layout (src/cpu/ivf.rs:25-48, 257-294; src/quantised/ivf_bf16.rs,
 * src/quantised/ivf_sq8.rs): vectors in list order in the index dtype (f32 / bf16 bit patterns /
 * int8 codes), per-vector norms (f32 for F32+BF16 cosine, int32 sum-of-squares for SQ8 cosine;
 * NULL for L2), f32 centroids (+ norms for cosine F32/BF16; NULL otherwise), CSR offsets
 * [nlist+1], original_ids [n] (list order -> original row), SQ8 scales [dim].
 * Sharding: `list_begin..list_end` is the range of lists whose vectors this handle stores
 * (`vectors`, `norms`, `original_ids` then hold only rows offsets[list_begin]..offsets[list_end]);
 * offsets always describe the whole index so probe expansion sees global list sizes.
 * Pass 0, nlist for an unsharded index. */
int annb_ivf_create(annb_index** out, const void* vectors, const void* norms, const float* centroids,
                    const float* centroid_norms, const uint64_t* offsets, const uint64_t* original_ids,
                    uint64_t n, uint32_t dim, uint32_t nlist, int dtype, int metric, const float* sq8_scales,
                    uint32_t list_begin, uint32_t list_end, int device);

/* Replaces IvfIndexGpu::query_batch (src/gpu/ivf_gpu.rs:466-503) / query_ivf_index_gpu
 * (src/lib.rs:2949) and IvfIndex / IvfIndexBf16 / IvfSq8Index::query
 * (src/cpu/ivf.rs:337-390, src/quantised/ivf_bf16.rs:277-332, src/quantised/ivf_sq8.rs:303-359).
 * nprobe == 0 means None -> max(1, floor(sqrt(nlist))); nprobe is a floor: probing expands in
 * centroid-rank order until k vectors are reachable (src/utils/k_means_utils.rs:3007-3029). */
int annb_ivf_search(const annb_index* index, const float* queries, uint64_t nq, uint32_t dim, uint32_t k,
                    uint32_t nprobe, uint64_t* out_ids, float* out_dist, uint32_t* out_counts);

/* Replaces IvfIndexGpu::generate_knn (src/gpu/ivf_gpu.rs:520-583) and the generate_knn of the
 * CPU / BF16 / SQ8 IVF indices.  Query set = stored rows in list order positions
 * [pos_begin, pos_end); output row j belongs to internal position pos_begin + j, unless
 * `scatter_to_original` is non-zero, in which case outputs are [n * k] and row r belongs to
 * original id r (src/cpu/ivf.rs:476-486).  Unsharded indices only. */
int annb_ivf_search_self(const annb_index* index, uint64_t pos_begin, uint64_t pos_end, uint32_t k,
                         uint32_t nprobe, int scatter_to_original, uint64_t* out_ids, float* out_dist,
                         uint32_t* out_counts);

int annb_ivf_search_dev(const annb_index* index, const float* d_queries, uint64_t nq, uint32_t dim, uint32_t k,
                        uint32_t nprobe, uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts,
                        void* stream);

/* The two halves of annb_ivf_search_dev, for sharded deployments: every rank holds the replicated centroid
 * table but only needs to rank the centroids for its slice of the query batch; the probe lists are exchanged
 * (all-gather of nq * probe_pitch cell ids) and every rank then scans its own lists for the whole batch.
 *   route : centroid ranking + probe expansion (src/cpu/ivf.rs:349-365) -> d_probes [nq * probe_pitch] (cell ids in
 *           rank order, UINT32_MAX padding), d_n_probes [nq].  ANNB_ERR_UNSUPPORTED if a query's expanded probe set
 *           does not fit probe_pitch (the caller then routes with a larger pitch or uses annb_ivf_search_dev).
 *   search_probes : list scan + top-k for the given probe lists (the same scan kernels as annb_ivf_search_dev). */
int annb_ivf_route_dev(const annb_index* index, const float* d_queries, uint64_t nq, uint32_t dim, uint32_t k, uint32_t nprobe,
                       uint32_t* d_probes, uint32_t* d_n_probes, uint32_t probe_pitch, void* stream);
int annb_ivf_search_probes_dev(const annb_index* index, const float* d_queries, uint64_t nq, uint32_t dim, uint32_t k, uint32_t nprobe,
                               const uint32_t* d_probes, const uint32_t* d_n_probes, uint32_t probe_pitch, uint64_t* d_out_ids,
                               float* d_out_dist, uint32_t* d_out_counts, void* stream);

/* ---------------------------------------------------------------- shared -- */

/* Multi-GPU exchange step: merges `parts` per-shard results (each [nq * k], laid out
 * [part][query][k], padded as described above) into one [nq * k] result under the
 * (distance, id) order.  Device pointers; runs on `stream`.  The all-gather that fills
 * d_part_* is NCCL's (torch.distributed / ncclAllGather) -- see DESIGN.md. */
int annb_merge_topk_dev(const uint64_t* d_part_ids, const float* d_part_dist, uint32_t parts, uint64_t nq,
                        uint32_t k, uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts,
                        void* stream);

/* The same exchange with one buffer per shard: shard p's [nq * k] ids start at d_parts + p * part_stride_bytes, its
 * [nq * k] distances dist_offset_bytes further on, so ONE all-gather (or one peer copy per shard) moves ids and distances
 * together.  Ties on distance keep the shards' own order, lower shard first: every shard returns its rows in the index's
 * own order -- (distance, row) flat, (distance, list position) IVF -- and shards own ascending disjoint row / list
 * ranges, so with the shards passed in that order the merged rows are bit-identical to the unsharded index's
 * (annb_merge_topk_dev orders ties by id instead, which differs from the reference for IVF lists). */
int annb_merge_shards_dev(const void* d_parts, uint64_t part_stride_bytes, uint64_t dist_offset_bytes, uint32_t parts,
                          uint64_t nq, uint32_t k, uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts,
                          void* stream);

/* Shard-mode searches: what a sharded deployment calls instead of annb_flat_search_dev / annb_ivf_search_probes_dev.  A
 * shard cannot know the k-th distance of the merged result, and being exact about its own k-th neighbour is wasted work
 * when that neighbour is far from the global top-k (a shard that holds only distant lists of a query sees densely
 * packed, uncertifiable candidates).  So these calls never recompute anything locally: next to the shard's k rows they
 * report d_out_bound[q], the distance below which no row of the shard that was NOT re-ranked exactly can lie (+inf where
 * every candidate was re-ranked).  After the exchange and the merge,
 *   annb_shard_check_dev  lists (inside the handle) the queries whose merged k-th distance is not strictly below this
 *                         shard's bound -- only those could be missing a row of this shard -- and returns their number
 *                         (one 4-byte read-back); and, if any shard reported any,
 *   annb_shard_refine_dev recomputes exactly those queries on the exact kernels and overwrites their rows in the
 *                         shard's result block, after which the exchange and the merge are repeated.
 * With every bound above the merged k-th distance the result is provably the unsharded one.  At most 16384 queries per
 * call for check / refine. */
int annb_flat_search_shard_dev(const annb_index* index, const float* d_queries, uint64_t nq, uint32_t dim, uint32_t k,
                               uint64_t* d_out_ids, float* d_out_dist, float* d_out_bound, void* stream);
int annb_ivf_search_probes_shard_dev(const annb_index* index, const float* d_queries, uint64_t nq, uint32_t dim, uint32_t k,
                                     uint32_t nprobe, const uint32_t* d_probes, const uint32_t* d_n_probes,
                                     uint32_t probe_pitch, uint64_t* d_out_ids, float* d_out_dist, float* d_out_bound,
                                     void* stream);
int annb_shard_check_dev(annb_index* index, const float* d_bound, const float* d_merged_dist, uint64_t nq, uint32_t k,
                         uint32_t* out_count, void* stream);
/* annb_shard_check_dev for deployments whose exchange already carries the bounds: every shard appends its [nq] bounds to
 * its result block (bound_offset_bytes into the block), and after the all-gather each rank tests ALL shards' bounds against
 * the merged rows -- *out_any (some shard has to refine) is then the same on every rank without another collective;
 * *out_mine is the number of queries listed for this shard (my_part).  The int32 that follows a shard's bounds in its block
 * is the shard's status word: bit 1 of *out_any reports that some shard wrote a non-zero status (its search failed), so all
 * ranks can give up on the step together. */
int annb_shard_check_gathered_dev(annb_index* index, const void* d_parts, uint64_t part_stride_bytes, uint64_t bound_offset_bytes,
                                  uint32_t parts, uint32_t my_part, const float* d_merged_dist, uint64_t nq, uint32_t k,
                                  uint32_t* out_mine, uint32_t* out_any, void* stream);

/* annb_shard_check_gathered_dev without the read-back: the verdict words are copied into the caller's pinned host buffer
 * h_verdict[2] on `stream` and the call returns at once; h_verdict[0] = this shard's queries to refine, h_verdict[1] = bit 0:
 * some shard has to refine, bit 1: some shard's call failed -- valid once an event recorded behind the call has completed.  A
 * serving loop can enqueue the next batch before the previous verdict is known; the list of queries to refine lives in the handle
 * only until its next search, so a deferred refine repeats the synchronous check first. */
int annb_shard_check_gathered_async_dev(annb_index* index, const void* d_parts, uint64_t part_stride_bytes, uint64_t bound_offset_bytes,
                                        uint32_t parts, uint32_t my_part, const float* d_merged_dist, uint64_t nq, uint32_t k,
                                        uint32_t* h_verdict, void* stream);

/* annb_merge_shards_dev and annb_shard_check_gathered_async_dev in ONE pass over the gathered blocks (the deferred step of a serving
 * loop: one kernel, one 8-byte copy to h_verdict[2]); same merged rows, same verdict words.  The handle keeps no list of queries to
 * refine from this call: a verdict that asks for a refine is followed by the synchronous annb_shard_check_gathered_dev.  Replaces the
 * merge + host-side bookkeeping that follows the per-shard queries of src/lib.rs:2842, 2949 in a sharded deployment. */
int annb_merge_check_shards_async_dev(annb_index* index, const void* d_parts, uint64_t part_stride_bytes, uint64_t dist_offset_bytes,
                                      uint64_t bound_offset_bytes, uint32_t parts, uint32_t my_part, uint64_t nq, uint32_t k,
                                      uint64_t* d_out_ids, float* d_out_dist, uint32_t* h_verdict, void* stream);

int annb_shard_refine_dev(annb_index* index, const float* d_queries, uint64_t nq, uint32_t dim, uint32_t k, uint32_t nprobe,
                          const uint32_t* d_probes, const uint32_t* d_n_probes, uint32_t probe_pitch, uint64_t* d_ids,
                          float* d_dist, void* stream);

/* ------------------------------------------------------- several devices -- */

/* One index over several GPUs of one box, driven from one process: the `device` argument of
 * build_exhaustive_index_gpu / build_ivf_index_gpu (src/lib.rs:2813, 2913) becomes a device list.  Rows (flat) or
 * inverted lists (IVF: contiguous list ranges balanced by vector count; centroid table and offsets replicated) are
 * sharded over `devices`; the returned handle is accepted by annb_flat_search, annb_flat_search_self, annb_ivf_search,
 * annb_flat_search_dev / annb_ivf_search_dev (buffers on devices[0]; these two then run synchronously),
 * annb_index_get_info / set_option / get_stat (options go to every shard, counters add up) and annb_destroy.
 * Per batch every device receives the queries straight from the caller's buffer, IVF devices rank the centroids for
 * their slice of the batch and exchange the probe lists by peer copies, every device searches its shard and stores its
 * per-shard top-k into its slot on devices[0] over NVLink, devices[0] merges (annb_merge_shards_dev's order: results are
 * bit-identical to the single-device index) and the result leaves with one device -> host copy.  Arguments as for
 * annb_flat_create / annb_ivf_create (SQ8 flat shards share one codebook trained on all rows). */
int annb_flat_create_multi(annb_index** out, const float* data, uint64_t n, uint32_t dim, int dtype, int metric,
                           const int* devices, int n_devices);
int annb_ivf_create_multi(annb_index** out, const void* vectors, const void* norms, const float* centroids,
                          const float* centroid_norms, const uint64_t* offsets, const uint64_t* original_ids,
                          uint64_t n, uint32_t dim, uint32_t nlist, int dtype, int metric, const float* sq8_scales,
                          const int* devices, int n_devices);
/* Number of per-device shards behind a handle (1 for an ordinary handle). */
int annb_index_shard_count(const annb_index* index, uint32_t* out);

/* KnnValidation::validate_index (src/utils/mod.rs:210-242; implemented for the CPU IvfIndex only, src/cpu/ivf.rs:496-523):
 * recall@k of the index against an exhaustive search over its own vectors, on the stored vectors at `positions` (internal
 * list-order positions in [0, n), drawn by the caller -- the reference draws them with rand's StdRng, which stays on the Rust
 * side).  nprobe = 0: the index default, as validate_index queries with None.  *out_recall = mean |approx ∩ true| / k.
 * Unsharded single-device f32 IVF handles only (ANNB_ERR_UNSUPPORTED otherwise). */
int annb_ivf_validate(const annb_index* index, const uint64_t* positions, uint64_t n_samples, uint32_t k, uint32_t nprobe,
                      double* out_recall);

/* Index facts: ExhaustiveIndexGpu::memory_usage_bytes (src/gpu/exhaustive_gpu.rs:205-209),
 * IvfIndexGpu::memory_usage_bytes (src/gpu/ivf_gpu.rs:590-604). */
typedef struct annb_index_info {
    uint64_t n;          /* vectors stored by this handle            */
    uint64_t n_total;    /* vectors of the whole (unsharded) index   */
    uint32_t dim;
    uint32_t nlist;      /* 0 for a flat index                       */
    int32_t dtype;
    int32_t metric;
    int32_t device;
    int32_t is_ivf;
    uint64_t device_bytes; /* HBM held by the handle                 */
    uint64_t host_bytes;
} annb_index_info;
int annb_index_get_info(const annb_index* index, annb_index_info* out);

/* Tunables / instrumentation.
 *   set_option: "path" (annb_path), "tc_candidates" (k' of the tensor-core pre-selection),
 *               "db_splits" (flat: database splits per query tile, 0 = auto),
 *               "scan_parts" (IVF: partial scans per query, 0 = auto),
 *               "ivf_list_major" (IVF list scan: -1 auto, 0 query-major streaming kernel, 1 list-major batched kernel),
 *               "ivf_task_order" (tensor-core IVF scan: 1 = (list x query group) tasks are handed out longest list first, 0 = in list order),
 *               "time_kernels" (1 = bracket the dominant kernel of every search with CUDA events on its stream),
 *               "tc_ts" (tensor paths: 1 = query operand resident in TMEM), "tc_bf16_hybrid" (flat BF16 index, f32 queries: 1 = third
 *               query term multiplied from shared memory instead of TMEM; experiment, default 0), "ivf_fast_probe" (0/1/2),
 *               "ivf_tc_coarse" (1 = rank the centroids on the tensor cores when nlist >= 512, 0 = CUDA-core ranking only),
 *               "cert_eps_log2" (error bound assumed by the coverage certificate of the tensor paths: negative = 2^value, 0 = certificate off,
 *               1 = derived per kernel from its MMA count, the default -- DESIGN.md section 3), "async_dev" (see Conventions),
 *               "cert_fallback" (1 = queries that fail the certificate are recomputed on the exact CUDA-core path),
 *               "tc_wide_k" (flat tensor path: 1 = serve 24 < k <= 256 -- the reference's radix-select range, src/gpu/topk_gpu.rs:95 --
 *               from the union of interleaved k' = 32 lists, and let a handle whose batches fail the certificate switch to that mode;
 *               0 = such k go to the CUDA-core path), "tc_strided" (1 = interleave the splits' tiles over the database also for k <= 24),
 *               "tc_f32_lo_smem" (f32 rows of <= 128 elements: 1 = lo query piece in shared memory, a third accumulator stage in TMEM),
 *               "tc_f32_fp16" (flat f32 index, rows of <= 256 elements: 1 (default) = 3xFP16 pre-selection -- rows scaled by powers of two,
 *               fp16 hi + lo pieces, the three product terms of 3xTF32 at twice the elements per MMA; 0 = 3xTF32.  Changing it rebuilds
 *               the handle's tensor-core operand copy)
 *   get_stat  : "kernel_launches" (cumulative), "scanned_vectors" (IVF, last call: sum of probed
 *               list lengths), "scanned_vectors_local" (the part of it that lies in this handle's own lists),
 *               "probed_lists" (last call), "last_path" (annb_path actually used),
 *               "coarse_path" (IVF, last call: 0 exact dense centroid ranking, 1 fused CUDA-core select, 2 tensor cores),
 *               "uncertified" (tensor path, last call: queries that failed the coverage certificate),
 *               "fallback_queries" (cumulative: queries recomputed on the exact path),
 *               "cert_eps_bits" (f32 bit pattern of the error bound the last tensor-path certificate assumed),
 *               "tc_kind" (flat: operand form of the tensor path: -1 none, 0 3xTF32, 1 bf16 terms, 2 int8, 3 3xFP16),
 *               "tc_escalated" (flat: 1 once a batch left more than 2 % of its queries uncertified -- later batches run in wide-k mode),
 *               "dominant_kernel_ns" / "dominant_kernel_launches" (with "time_kernels": summed device time and count
 *               of the dominant kernel -- flat distance+select kernel or IVF list-scan kernel -- since the option was set) */
int annb_index_set_option(annb_index* index, const char* key, int64_t value);
int annb_index_get_stat(const annb_index* index, const char* key, int64_t* out);

/* Diagnostics (tests only): with option "tc_debug" = 1 the first CTA of the tensor-core flat kernel dumps the
 * 128 x 128 selection values v = fma(q.x, a, b) of its first tile; this copies them to host_out[128 * 128].
 * IVF handle: the first 64 KiB of the packed (value, row) candidate lists of the last tensor-core scan. */
int annb_debug_fetch_tile(annb_index* index, float* host_out);
/* ... and its role wait-cycle counters {mma_total, producer_wait_empty, mma_wait_full, mma_wait_tmem_empty,
 * epilogue_wait_tmem_full, epilogue_slow_path, tiles, 0} of the last tensor-path launch.
 * IVF handle: {total, schedule, query gather, epilogue wait-tmem-full, mma wait-queries, mma wait-data,
 * mma wait-tmem-empty, tasks << 32 | tiles} of CTA 0 of the last tensor-core scan. */
int annb_debug_fetch_cycles(annb_index* index, uint64_t* host_out8);
/* The queries (batch-relative numbers) of the last tensor-path batch that failed the coverage certificate: *out_count
 * receives their number, host_out the first min(count, capacity) of them.  With option "cert_fallback" = 0 these are the
 * only rows of the result that may differ from the exact answer. */
int annb_debug_fetch_uncertified(annb_index* index, uint32_t* host_out, uint32_t capacity, uint32_t* out_count);

void annb_destroy(annb_index* index);

#ifdef __cplusplus
}
#endif
#endif /* ANNB200_H */
