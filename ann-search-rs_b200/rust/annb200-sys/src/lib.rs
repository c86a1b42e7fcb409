//! Rust side of the drop-in boundary: `extern "C"` declarations of include/annb200.h plus safe wrappers that keep
//! the reference's `build_*` / `query_*` signatures (faer `MatRef` in, `(indices, distances)` out).
//!
//! Written against ann-search-rs v0.5.2; every item cites the reference item it replaces.  This file is not
//! compiled in the build image (no Rust toolchain) -- the C++ and Python mirrors exercise the same C ABI there.
#![allow(clippy::too_many_arguments)]
use faer::MatRef;
use std::ffi::{c_char, c_int, c_void, CStr, CString};

#[repr(C)]
pub struct annb_index {
    _private: [u8; 0],
}

pub const ANNB_F32: c_int = 0;
pub const ANNB_BF16: c_int = 1;
pub const ANNB_SQ8: c_int = 2;
pub const ANNB_L2: c_int = 0;
pub const ANNB_COSINE: c_int = 1;
pub const ANNB_MANHATTAN: c_int = 2;

extern "C" {
    pub fn annb_last_error() -> *const c_char;
    pub fn annb_parse_metric(s: *const c_char) -> c_int;
    pub fn annb_flat_create(out: *mut *mut annb_index, data: *const f32, n: u64, dim: u32, dtype: c_int, metric: c_int,
                            sq8_scales: *const f32, id_base: u64, device: c_int) -> c_int;
    pub fn annb_flat_search(index: *const annb_index, queries: *const f32, nq: u64, dim: u32, k: u32, out_ids: *mut u64,
                            out_dist: *mut f32, out_counts: *mut u32) -> c_int;
    pub fn annb_flat_search_self(index: *const annb_index, row_begin: u64, row_end: u64, k: u32, out_ids: *mut u64,
                                 out_dist: *mut f32, out_counts: *mut u32) -> c_int;
    pub fn annb_ivf_assign(data: *const f32, n: u64, dim: u32, centroids: *const f32, centroid_norms: *const f32, nlist: u32,
                           metric: c_int, out_assign: *mut u32, device: c_int) -> c_int;
    /// `matrix_to_flat` (src/utils/mod.rs:44-68) on the device; `mat` / `out_rowmajor` host or device memory, strides in elements.
    pub fn annb_matrix_to_flat(mat: *const f32, nrows: u64, ncols: u32, row_stride: i64, col_stride: i64, out_rowmajor: *mut f32,
                               device: c_int) -> c_int;
    /// bincode "standard" varints of the crate's `Vec<usize>` fields (host only; the Rust side has bincode itself).
    pub fn annb_varint_encode_u64(values: *const u64, count: u64, out: *mut u8, out_capacity: u64) -> i64;
    pub fn annb_varint_decode_u64(buf: *const u8, len: u64, count: u64, out: *mut u64) -> i64;
    /// Diagnostic: rows of this thread's last assign / Lloyd call that failed the tensor path's certificate and were redone exactly.
    pub fn annb_assign_last_redone() -> u64;
    pub fn annb_ivf_route_dev(index: *const annb_index, d_queries: *const f32, nq: u64, dim: u32, k: u32, nprobe: u32, d_probes: *mut u32,
                              d_n_probes: *mut u32, probe_pitch: u32, stream: *mut c_void) -> c_int;
    pub fn annb_ivf_search_probes_dev(index: *const annb_index, d_queries: *const f32, nq: u64, dim: u32, k: u32, nprobe: u32,
                                      d_probes: *const u32, d_n_probes: *const u32, probe_pitch: u32, d_out_ids: *mut u64,
                                      d_out_dist: *mut f32, d_out_counts: *mut u32, stream: *mut c_void) -> c_int;
    pub fn annb_kmeans_lloyd(data: *const f32, n: u64, dim: u32, centroids: *mut f32, nlist: u32, metric: c_int, max_iters: u32,
                             out_iters: *mut u32, device: c_int) -> c_int;
    pub fn annb_ivf_create(out: *mut *mut annb_index, vectors: *const c_void, norms: *const c_void, centroids: *const f32,
                           centroid_norms: *const f32, offsets: *const u64, original_ids: *const u64, n: u64, dim: u32,
                           nlist: u32, dtype: c_int, metric: c_int, sq8_scales: *const f32, list_begin: u32, list_end: u32,
                           device: c_int) -> c_int;
    pub fn annb_ivf_search(index: *const annb_index, queries: *const f32, nq: u64, dim: u32, k: u32, nprobe: u32,
                           out_ids: *mut u64, out_dist: *mut f32, out_counts: *mut u32) -> c_int;
    pub fn annb_ivf_search_self(index: *const annb_index, pos_begin: u64, pos_end: u64, k: u32, nprobe: u32,
                                scatter_to_original: c_int, out_ids: *mut u64, out_dist: *mut f32, out_counts: *mut u32) -> c_int;
    pub fn annb_destroy(index: *mut annb_index);
    pub fn annb_version() -> c_int;
    pub fn annb_device_count(out: *mut c_int) -> c_int;
    pub fn annb_flat_search_dev(index: *const annb_index, d_queries: *const f32, nq: u64, dim: u32, k: u32, d_out_ids: *mut u64,
                                d_out_dist: *mut f32, d_out_counts: *mut u32, stream: *mut c_void) -> c_int;
    pub fn annb_ivf_search_dev(index: *const annb_index, d_queries: *const f32, nq: u64, dim: u32, k: u32, nprobe: u32, d_out_ids: *mut u64,
                               d_out_dist: *mut f32, d_out_counts: *mut u32, stream: *mut c_void) -> c_int;
    /// Exact kNN-graph rows in the shape of `KnnGraphGpu` (src/gpu/nndescent_gpu.rs:2418-2446): self edge dropped, sentinel padded.
    pub fn annb_flat_knn_graph(index: *const annb_index, row_begin: u64, row_end: u64, k: u32, out_pid: *mut u64, out_dist: *mut f32,
                               out_counts: *mut u32) -> c_int;
    /// `parallel_lloyd` with the balancing hook of `KMeansTrainingParams::with_balancing` (src/utils/k_means_utils.rs:979-1030, 1572-1700).
    pub fn annb_kmeans_lloyd_balanced(data: *const f32, n: u64, dim: u32, centroids: *mut f32, nlist: u32, metric: c_int, max_iters: u32,
                                      balanced: c_int, seed: u64, out_iters: *mut u32, out_adjusted: *mut u64, device: c_int) -> c_int;
    pub fn annb_merge_topk_dev(d_part_ids: *const u64, d_part_dist: *const f32, parts: u32, nq: u64, k: u32, d_out_ids: *mut u64,
                               d_out_dist: *mut f32, d_out_counts: *mut u32, stream: *mut c_void) -> c_int;
    /// One buffer per shard ([ids | distances]), ties in shard order = the unsharded order.
    pub fn annb_merge_shards_dev(d_parts: *const c_void, part_stride_bytes: u64, dist_offset_bytes: u64, parts: u32, nq: u64, k: u32,
                                 d_out_ids: *mut u64, d_out_dist: *mut f32, d_out_counts: *mut u32, stream: *mut c_void) -> c_int;
    /// One index over several GPUs of the box behind one handle (rows / inverted lists sharded, peer copies over NVLink).
    pub fn annb_flat_create_multi(out: *mut *mut annb_index, data: *const f32, n: u64, dim: u32, dtype: c_int, metric: c_int,
                                  devices: *const c_int, n_devices: c_int) -> c_int;
    pub fn annb_ivf_create_multi(out: *mut *mut annb_index, vectors: *const c_void, norms: *const c_void, centroids: *const f32,
                                 centroid_norms: *const f32, offsets: *const u64, original_ids: *const u64, n: u64, dim: u32, nlist: u32,
                                 dtype: c_int, metric: c_int, sq8_scales: *const f32, devices: *const c_int, n_devices: c_int) -> c_int;
    /// Shard-mode searches: a bound instead of a local certificate; tested against the merged rows afterwards.
    pub fn annb_flat_search_shard_dev(index: *const annb_index, d_queries: *const f32, nq: u64, dim: u32, k: u32, d_out_ids: *mut u64,
                                      d_out_dist: *mut f32, d_out_bound: *mut f32, stream: *mut c_void) -> c_int;
    pub fn annb_ivf_search_probes_shard_dev(index: *const annb_index, d_queries: *const f32, nq: u64, dim: u32, k: u32, nprobe: u32,
                                            d_probes: *const u32, d_n_probes: *const u32, probe_pitch: u32, d_out_ids: *mut u64,
                                            d_out_dist: *mut f32, d_out_bound: *mut f32, stream: *mut c_void) -> c_int;
    pub fn annb_shard_check_dev(index: *mut annb_index, d_bound: *const f32, d_merged_dist: *const f32, nq: u64, k: u32, out_count: *mut u32,
                                stream: *mut c_void) -> c_int;
    pub fn annb_shard_check_gathered_dev(index: *mut annb_index, d_parts: *const c_void, part_stride_bytes: u64, bound_offset_bytes: u64, parts: u32,
                                         my_part: u32, d_merged_dist: *const f32, nq: u64, k: u32, out_mine: *mut u32, out_any: *mut u32,
                                         stream: *mut c_void) -> c_int;
    pub fn annb_ivf_validate(index: *const annb_index, positions: *const u64, n_samples: u64, k: u32, nprobe: u32, out_recall: *mut f64) -> c_int;
    pub fn annb_shard_check_gathered_async_dev(index: *mut annb_index, d_parts: *const c_void, part_stride_bytes: u64, bound_offset_bytes: u64, parts: u32,
                                               my_part: u32, d_merged_dist: *const f32, nq: u64, k: u32, h_verdict: *mut u32, stream: *mut c_void) -> c_int;
    pub fn annb_merge_check_shards_async_dev(index: *mut annb_index, d_parts: *const c_void, part_stride_bytes: u64, dist_offset_bytes: u64,
                                             bound_offset_bytes: u64, parts: u32, my_part: u32, nq: u64, k: u32, d_out_ids: *mut u64,
                                             d_out_dist: *mut f32, h_verdict: *mut u32, stream: *mut c_void) -> c_int;
    pub fn annb_shard_refine_dev(index: *mut annb_index, d_queries: *const f32, nq: u64, dim: u32, k: u32, nprobe: u32, d_probes: *const u32,
                                 d_n_probes: *const u32, probe_pitch: u32, d_ids: *mut u64, d_dist: *mut f32, stream: *mut c_void) -> c_int;
    pub fn annb_index_shard_count(index: *const annb_index, out: *mut u32) -> c_int;
    pub fn annb_index_get_info(index: *const annb_index, out: *mut annb_index_info) -> c_int;
    pub fn annb_index_set_option(index: *mut annb_index, key: *const c_char, value: i64) -> c_int;
    pub fn annb_index_get_stat(index: *const annb_index, key: *const c_char, out: *mut i64) -> c_int;
    pub fn annb_debug_fetch_tile(index: *mut annb_index, host_out: *mut f32) -> c_int;
    pub fn annb_debug_fetch_cycles(index: *mut annb_index, host_out8: *mut u64) -> c_int;
    pub fn annb_debug_fetch_uncertified(index: *mut annb_index, host_out: *mut u32, capacity: u32, out_count: *mut u32) -> c_int;
}

/// `annb_index_info` of include/annb200.h (memory_usage_bytes of the reference's GPU indices, src/gpu/ivf_gpu.rs:590-604).
#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct annb_index_info {
    pub n: u64,
    pub n_total: u64,
    pub dim: u32,
    pub nlist: u32,
    pub dtype: i32,
    pub metric: i32,
    pub device: i32,
    pub is_ivf: i32,
    pub device_bytes: u64,
    pub host_bytes: u64,
}

/// Subset of `AnnSearchErrors` (src/errors.rs) reachable from this path; the status codes map 1:1.
#[derive(Debug, thiserror::Error)]
pub enum AnnSearchErrors {
    #[error("dimension mismatch: {0}")]
    DimensionMismatch(String),       // errors.rs:23
    #[error("distance not supported: {0}")]
    DistanceNotSupported(String),    // errors.rs:32
    #[error("too few samples for centroids: {0}")]
    TooFewSamplesForCentroids(String), // errors.rs:90
    #[error("CUDA error: {0}")]
    Cuda(String),                    // replaces CubeClServerError / CubeclUtils, errors.rs:212-223
    #[error("{0}")]
    Other(String),
}

fn check(status: c_int) -> Result<(), AnnSearchErrors> {
    if status == 0 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(annb_last_error()) }.to_string_lossy().into_owned();
    Err(match status {
        -1 => AnnSearchErrors::DimensionMismatch(msg),
        -2 => AnnSearchErrors::DistanceNotSupported(msg),
        -3 => AnnSearchErrors::TooFewSamplesForCentroids(msg),
        -5 => AnnSearchErrors::Cuda(msg),
        _ => AnnSearchErrors::Other(msg),
    })
}

pub type KnnOptionResult = Result<(Vec<Vec<usize>>, Option<Vec<Vec<f32>>>), AnnSearchErrors>;

/// matrix_to_flat (src/utils/mod.rs:44-68).
fn matrix_to_flat(mat: MatRef<f32>) -> (Vec<f32>, usize, usize) {
    let (n, dim) = (mat.nrows(), mat.ncols());
    let mut flat = Vec::with_capacity(n * dim);
    for i in 0..n {
        for j in 0..dim {
            flat.push(mat[(i, j)]);
        }
    }
    (flat, n, dim)
}

fn metric_or_default(dist_metric: &str) -> c_int {
    let c = CString::new(dist_metric).unwrap();
    let m = unsafe { annb_parse_metric(c.as_ptr()) };
    if m < 0 {
        println!("  Unknown distance metric '{dist_metric}', defaulting to Euclidean"); // src/lib.rs:274-277
        ANNB_L2
    } else {
        m
    }
}

/// Replaces `ExhaustiveIndexGpu<T, R>` (src/gpu/exhaustive_gpu.rs:17-33): the database lives on the B200.
pub struct ExhaustiveIndexB200 {
    handle: *mut annb_index,
    pub n: usize,
    pub dim: usize,
}
unsafe impl Send for ExhaustiveIndexB200 {}
unsafe impl Sync for ExhaustiveIndexB200 {} // searches lock internally
impl Drop for ExhaustiveIndexB200 {
    fn drop(&mut self) {
        unsafe { annb_destroy(self.handle) }
    }
}

fn unpack(ids: Vec<u64>, dist: Vec<f32>, cnt: Vec<u32>, k: usize, return_dist: bool) -> (Vec<Vec<usize>>, Option<Vec<Vec<f32>>>) {
    let indices = cnt.iter().enumerate().map(|(i, &c)| ids[i * k..i * k + c as usize].iter().map(|&v| v as usize).collect()).collect();
    let distances = return_dist.then(|| cnt.iter().enumerate().map(|(i, &c)| dist[i * k..i * k + c as usize].to_vec()).collect());
    (indices, distances)
}

fn build_flat(mat: MatRef<f32>, dist_metric: &str, dtype: c_int, device: i32) -> Result<ExhaustiveIndexB200, AnnSearchErrors> {
    let (flat, n, dim) = matrix_to_flat(mat);
    let mut handle = std::ptr::null_mut();
    check(unsafe { annb_flat_create(&mut handle, flat.as_ptr(), n as u64, dim as u32, dtype, metric_or_default(dist_metric),
                                    std::ptr::null(), 0, device) })?;
    Ok(ExhaustiveIndexB200 { handle, n, dim })
}

/// src/lib.rs:2813 `build_exhaustive_index_gpu` (device = CUDA ordinal instead of `R::Device`).
pub fn build_exhaustive_index_gpu(mat: MatRef<f32>, dist_metric: &str, device: i32) -> Result<ExhaustiveIndexB200, AnnSearchErrors> {
    build_flat(mat, dist_metric, ANNB_F32, device)
}
/// src/lib.rs:1702 `build_exhaustive_bf16_index`.
pub fn build_exhaustive_bf16_index(mat: MatRef<f32>, dist_metric: &str, device: i32) -> Result<ExhaustiveIndexB200, AnnSearchErrors> {
    build_flat(mat, dist_metric, ANNB_BF16, device)
}
/// src/lib.rs:1787 `build_exhaustive_sq8_index`.
pub fn build_exhaustive_sq8_index(mat: MatRef<f32>, dist_metric: &str, device: i32) -> Result<ExhaustiveIndexB200, AnnSearchErrors> {
    build_flat(mat, dist_metric, ANNB_SQ8, device)
}

/// src/lib.rs:2842 `query_exhaustive_index_gpu` (also serves the BF16 / SQ8 query functions, lib.rs:1733, 1818).
pub fn query_exhaustive_index_gpu(query_mat: MatRef<f32>, index: &ExhaustiveIndexB200, k: usize, return_dist: bool, _verbose: bool) -> KnnOptionResult {
    let (flat, nq, dim) = matrix_to_flat(query_mat);
    let (mut ids, mut dist, mut cnt) = (vec![0u64; nq * k], vec![0f32; if return_dist { nq * k } else { 0 }], vec![0u32; nq]);
    let dptr = if return_dist { dist.as_mut_ptr() } else { std::ptr::null_mut() };
    check(unsafe { annb_flat_search(index.handle, flat.as_ptr(), nq as u64, dim as u32, k as u32, ids.as_mut_ptr(), dptr, cnt.as_mut_ptr()) })?;
    Ok(unpack(ids, dist, cnt, k, return_dist))
}

/// src/lib.rs:2877 `query_exhaustive_index_gpu_self`.
pub fn query_exhaustive_index_gpu_self(index: &ExhaustiveIndexB200, k: usize, return_dist: bool, _verbose: bool) -> KnnOptionResult {
    let n = index.n;
    let (mut ids, mut dist, mut cnt) = (vec![0u64; n * k], vec![0f32; if return_dist { n * k } else { 0 }], vec![0u32; n]);
    let dptr = if return_dist { dist.as_mut_ptr() } else { std::ptr::null_mut() };
    check(unsafe { annb_flat_search_self(index.handle, 0, n as u64, k as u32, ids.as_mut_ptr(), dptr, cnt.as_mut_ptr()) })?;
    Ok(unpack(ids, dist, cnt, k, return_dist))
}

/// Replaces `IvfIndexGpu<T, R>` (src/gpu/ivf_gpu.rs:153-181).  Built from a CPU-side `IvfIndex` / `IvfIndexBf16` /
/// `IvfSq8Index` (the crate keeps its own k-means; coarse assignment can call `annb_ivf_assign`).
pub struct IvfIndexB200 {
    handle: *mut annb_index,
    pub n: usize,
    pub dim: usize,
    pub nlist: usize,
}
unsafe impl Send for IvfIndexB200 {}
unsafe impl Sync for IvfIndexB200 {}
impl Drop for IvfIndexB200 {
    fn drop(&mut self) {
        unsafe { annb_destroy(self.handle) }
    }
}

impl IvfIndexB200 {
    /// Fields as in `IvfIndex<T>` after `optimise_memory_layout` (src/cpu/ivf.rs:25-48, 257-294).
    pub fn from_parts_f32(vectors_flat: &[f32], norms: &[f32], centroids: &[f32], centroids_norm: &[f32], offsets: &[usize],
                          original_ids: &[usize], dim: usize, metric: c_int, device: i32) -> Result<Self, AnnSearchErrors> {
        let nlist = offsets.len() - 1;
        let n = original_ids.len();
        let off: Vec<u64> = offsets.iter().map(|&v| v as u64).collect();
        let ids: Vec<u64> = original_ids.iter().map(|&v| v as u64).collect();
        let mut handle = std::ptr::null_mut();
        let np = if norms.is_empty() { std::ptr::null() } else { norms.as_ptr() as *const c_void };
        let cp = if centroids_norm.is_empty() { std::ptr::null() } else { centroids_norm.as_ptr() };
        check(unsafe { annb_ivf_create(&mut handle, vectors_flat.as_ptr() as *const c_void, np, centroids.as_ptr(), cp, off.as_ptr(),
                                       ids.as_ptr(), n as u64, dim as u32, nlist as u32, ANNB_F32, metric, std::ptr::null(), 0,
                                       nlist as u32, device) })?;
        Ok(Self { handle, n, dim, nlist })
    }

    /// `KnnValidation::validate_index` (src/utils/mod.rs:210-242): recall@k against an exhaustive search over the index's own
    /// vectors.  The caller draws the sample positions (`rng.random_range(0..n)` with the crate's StdRng, as the reference does).
    pub fn validate_index(&self, k: usize, positions: &[usize]) -> Result<f64, AnnSearchErrors> {
        let pos: Vec<u64> = positions.iter().map(|&v| v as u64).collect();
        let mut recall = 0f64;
        check(unsafe { annb_ivf_validate(self.handle, pos.as_ptr(), pos.len() as u64, k as u32, 0, &mut recall) })?;
        Ok(recall)
    }
}

/// src/lib.rs:2949 `query_ivf_index_gpu` (`nquery`, the reference's batch size, is accepted and ignored).
pub fn query_ivf_index_gpu(query_mat: MatRef<f32>, index: &IvfIndexB200, k: usize, nprobe: Option<usize>, _nquery: Option<usize>,
                           return_dist: bool, _verbose: bool) -> KnnOptionResult {
    let (flat, nq, dim) = matrix_to_flat(query_mat);
    let (mut ids, mut dist, mut cnt) = (vec![0u64; nq * k], vec![0f32; if return_dist { nq * k } else { 0 }], vec![0u32; nq]);
    let dptr = if return_dist { dist.as_mut_ptr() } else { std::ptr::null_mut() };
    check(unsafe { annb_ivf_search(index.handle, flat.as_ptr(), nq as u64, dim as u32, k as u32, nprobe.unwrap_or(0) as u32,
                                   ids.as_mut_ptr(), dptr, cnt.as_mut_ptr()) })?;
    Ok(unpack(ids, dist, cnt, k, return_dist))
}

/// src/lib.rs:2989 `query_ivf_index_gpu_self`.
pub fn query_ivf_index_gpu_self(index: &IvfIndexB200, k: usize, nprobe: Option<usize>, _nquery: Option<usize>, return_dist: bool,
                                _verbose: bool) -> KnnOptionResult {
    let n = index.n;
    let (mut ids, mut dist, mut cnt) = (vec![u64::MAX; n * k], vec![f32::INFINITY; if return_dist { n * k } else { 0 }], vec![0u32; n]);
    let dptr = if return_dist { dist.as_mut_ptr() } else { std::ptr::null_mut() };
    check(unsafe { annb_ivf_search_self(index.handle, 0, n as u64, k as u32, nprobe.unwrap_or(0) as u32, 1, ids.as_mut_ptr(), dptr,
                                        cnt.as_mut_ptr()) })?;
    Ok(unpack(ids, dist, cnt, k, return_dist))
}


// ------------------------------------------------------------------------------------------------------------------
// IVF builds (src/lib.rs:2913 build_ivf_index_gpu, :2100 build_ivf_bf16_index, :2196 build_ivf_sq8_index).
// The steps of IvfIndex::build (src/cpu/ivf.rs:145-249) / IvfIndexBf16::build (src/quantised/ivf_bf16.rs:150-254) /
// IvfSq8Index::build (src/quantised/ivf_sq8.rs:158-284): training subsample -> centroids (device Lloyd) -> coarse
// assignment (device) -> CSR layout and list-order permutation (host, integer work) -> quantisation -> resident index.
// Inside the crate the subsample and the seeding come from its own `sample_vectors` / `train_centroids`
// (src/utils/k_means_utils.rs:2771, 3047); this standalone file draws with SplitMix64 instead of rand's StdRng.
// ------------------------------------------------------------------------------------------------------------------

/// Mirror of `KMeansTrainingParams` (src/utils/k_means_utils.rs:286-345) as far as the device loop uses it.
#[derive(Debug, Clone, Copy)]
pub struct KMeansTrainingParams {
    pub iters: usize,
    pub balanced: bool,
}
impl Default for KMeansTrainingParams {
    fn default() -> Self {
        Self { iters: 30, balanced: false }
    }
}

struct SplitMix64(u64);
impl SplitMix64 {
    fn next(&mut self) -> u64 {
        self.0 = self.0.wrapping_add(0x9E3779B97F4A7C15);
        let mut z = self.0;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
        z ^ (z >> 31)
    }
}

/// First `m` entries of a Fisher-Yates shuffle of 0..n (stand-in for `indices.shuffle(&mut StdRng)`).
fn sample_rows(n: usize, m: usize, seed: u64) -> Vec<usize> {
    let mut idx: Vec<usize> = (0..n).collect();
    let mut rng = SplitMix64(seed);
    for i in 0..m.min(n) {
        let j = i + (rng.next() % (n - i) as u64) as usize;
        idx.swap(i, j);
    }
    idx.truncate(m.min(n));
    idx
}

/// build_csr_layout (src/utils/k_means_utils.rs:2955-2980): stable counting sort -> (new_to_old, offsets).
pub fn build_csr_layout(assign: &[u32], nlist: usize) -> (Vec<usize>, Vec<usize>) {
    let mut offsets = vec![0usize; nlist + 1];
    for &a in assign {
        offsets[a as usize + 1] += 1;
    }
    for c in 0..nlist {
        offsets[c + 1] += offsets[c];
    }
    let mut cursor = offsets.clone();
    let mut order = vec![0usize; assign.len()];
    for (i, &a) in assign.iter().enumerate() {
        order[cursor[a as usize]] = i;
        cursor[a as usize] += 1;
    }
    (order, offsets)
}

/// encode_bf16_quantisation (src/quantised/quantisers.rs:31-38): round to nearest even, as `half::bf16::from_f32`.
fn f32_to_bf16_bits(x: f32) -> u16 {
    let b = x.to_bits();
    if (b & 0x7FFF_FFFF) > 0x7F80_0000 {
        return ((b >> 16) | 0x40) as u16;
    }
    let round = ((b & 0x8000) != 0) && ((b & 0x1_7FFF) != 0);
    ((b >> 16) + round as u32) as u16
}

/// ScalarQuantiser::{train, encode} (src/quantised/quantisers.rs:123-165).
fn sq8_train(rows: &[f32], dim: usize) -> Vec<f32> {
    let mut mx = vec![0f32; dim];
    for r in rows.chunks_exact(dim) {
        for d in 0..dim {
            mx[d] = mx[d].max(r[d].abs());
        }
    }
    mx.iter().map(|&m| if m <= 0.0 { 1.0 } else { m / 128.0 }).collect()
}
fn sq8_encode(v: f32, scale: f32) -> i8 {
    let s = v / scale;
    if s.is_nan() {
        return 0;
    }
    (s + 0.5 * s.signum()).clamp(-128.0, 127.0) as i8
}

fn l2_norm(v: &[f32]) -> f32 {
    // calculate_l2_norm (src/utils/dist.rs:2339-2360): 8 lane accumulators, wide's reduce tree, scalar tail
    let mut acc = [0f32; 8];
    let chunks = v.len() / 8;
    for c in 0..chunks {
        for j in 0..8 {
            let x = v[c * 8 + j];
            acc[j] += x * x;
        }
    }
    let (s0, s1, s2, s3) = (acc[0] + acc[4], acc[1] + acc[5], acc[2] + acc[6], acc[3] + acc[7]);
    let mut sum = (s0 + s2) + (s1 + s3);
    for &x in &v[chunks * 8..] {
        sum += x * x;
    }
    sum.sqrt()
}

fn build_ivf(mat: MatRef<f32>, nlist: Option<usize>, k_means_params: Option<KMeansTrainingParams>, dist_metric: &str, seed: usize,
             verbose: bool, dtype: c_int, devices: &[i32]) -> Result<IvfIndexB200, AnnSearchErrors> {
    let metric = metric_or_default(dist_metric);
    if metric == ANNB_MANHATTAN {
        return Err(AnnSearchErrors::DistanceNotSupported("Manhattan".into())); // src/cpu/ivf.rs:153-155
    }
    let (mut flat, n, dim) = matrix_to_flat(mat);
    let nlist = nlist.unwrap_or(((n as f32).sqrt() as usize).max(1)); // src/cpu/ivf.rs:172
    let params = k_means_params.unwrap_or_default();
    let dev0 = devices.first().copied().unwrap_or(0);
    let cosine = metric == ANNB_COSINE;
    // norms of the un-rounded rows (f32 / bf16 cosine); the SQ8 index normalises data and centroids instead (ivf_sq8.rs:169-205)
    let mut norms: Vec<f32> = Vec::new();
    if cosine && dtype != ANNB_SQ8 {
        norms = flat.chunks_exact(dim).map(l2_norm).collect();
    }
    if cosine && dtype == ANNB_SQ8 {
        for r in flat.chunks_exact_mut(dim) {
            let nr = l2_norm(r);
            if nr > 0.0 {
                r.iter_mut().for_each(|x| *x /= nr);
            }
        }
    }
    let n_train = (256 * nlist).min(250_000).min(n).max(1); // src/cpu/ivf.rs:174
    let rows = sample_rows(n, n_train, seed as u64);
    let mut train = Vec::with_capacity(n_train * dim);
    for &r in &rows {
        train.extend_from_slice(&flat[r * dim..(r + 1) * dim]);
    }
    if n_train < nlist {
        return Err(AnnSearchErrors::TooFewSamplesForCentroids(format!("{n_train} training samples for {nlist} centroids")));
    }
    if verbose {
        println!("  Generating IVF index with {nlist} Voronoi cells.");
    }
    // seeding: nlist distinct training rows (fast_random_init, the reference's choice for nlist > 200; `train` is already a shuffle)
    let mut centroids: Vec<f32> = train[..nlist * dim].to_vec();
    let mut iters = 0u32;
    check(unsafe { annb_kmeans_lloyd_balanced(train.as_ptr(), n_train as u64, dim as u32, centroids.as_mut_ptr(), nlist as u32, metric,
                                              params.iters as u32, params.balanced as c_int, seed as u64, &mut iters, std::ptr::null_mut(), dev0) })?;
    let mut scales: Vec<f32> = Vec::new();
    let mut cnorms: Vec<f32> = Vec::new();
    if dtype == ANNB_SQ8 {
        if cosine {
            for c in centroids.chunks_exact_mut(dim) {
                let nr = l2_norm(c);
                if nr > 0.0 {
                    c.iter_mut().for_each(|x| *x /= nr);
                }
            }
        }
        scales = sq8_train(&train, dim); // codebook from the training sample only (ivf_sq8.rs:211)
        cnorms = vec![1.0; nlist]; // direct_assign with unit norms (ivf_sq8.rs:214-215)
    } else if cosine {
        // sequential fold (src/cpu/ivf.rs:193-206)
        cnorms = centroids.chunks_exact(dim).map(|c| c.iter().fold(0f32, |s, &x| s + x * x).sqrt()).collect();
    }
    let mut assign = vec![0u32; n];
    let cn_ptr = if cnorms.is_empty() { std::ptr::null() } else { cnorms.as_ptr() };
    check(unsafe { annb_ivf_assign(flat.as_ptr(), n as u64, dim as u32, centroids.as_ptr(), cn_ptr, nlist as u32, metric, assign.as_mut_ptr(), dev0) })?;
    let (order, offsets) = build_csr_layout(&assign, nlist);
    // list-order permutation + quantisation
    let mut vec_f32: Vec<f32> = Vec::new();
    let mut vec_bf16: Vec<u16> = Vec::new();
    let mut vec_i8: Vec<i8> = Vec::new();
    let mut norms_lo: Vec<f32> = Vec::new();
    let mut norms_i: Vec<i32> = Vec::new();
    for &old in &order {
        let r = &flat[old * dim..(old + 1) * dim];
        match dtype {
            ANNB_F32 => vec_f32.extend_from_slice(r),
            ANNB_BF16 => vec_bf16.extend(r.iter().map(|&x| f32_to_bf16_bits(x))),
            _ => {
                let start = vec_i8.len();
                vec_i8.extend(r.iter().zip(&scales).map(|(&x, &s)| sq8_encode(x, s)));
                if cosine {
                    norms_i.push(vec_i8[start..].iter().map(|&c| c as i32 * c as i32).sum());
                }
            }
        }
        if !norms.is_empty() {
            norms_lo.push(norms[old]);
        }
    }
    let vptr: *const c_void = match dtype {
        ANNB_F32 => vec_f32.as_ptr() as *const c_void,
        ANNB_BF16 => vec_bf16.as_ptr() as *const c_void,
        _ => vec_i8.as_ptr() as *const c_void,
    };
    let nptr: *const c_void = if !norms_lo.is_empty() { norms_lo.as_ptr() as *const c_void } else if !norms_i.is_empty() { norms_i.as_ptr() as *const c_void } else { std::ptr::null() };
    let cnp = if cosine && dtype != ANNB_SQ8 { cnorms.as_ptr() } else { std::ptr::null() };
    let sp = if scales.is_empty() { std::ptr::null() } else { scales.as_ptr() };
    let off: Vec<u64> = offsets.iter().map(|&v| v as u64).collect();
    let ids: Vec<u64> = order.iter().map(|&v| v as u64).collect();
    let mut handle = std::ptr::null_mut();
    if devices.len() > 1 {
        check(unsafe { annb_ivf_create_multi(&mut handle, vptr, nptr, centroids.as_ptr(), cnp, off.as_ptr(), ids.as_ptr(), n as u64, dim as u32,
                                             nlist as u32, dtype, metric, sp, devices.as_ptr(), devices.len() as c_int) })?;
    } else {
        check(unsafe { annb_ivf_create(&mut handle, vptr, nptr, centroids.as_ptr(), cnp, off.as_ptr(), ids.as_ptr(), n as u64, dim as u32,
                                       nlist as u32, dtype, metric, sp, 0, nlist as u32, dev0) })?;
    }
    Ok(IvfIndexB200 { handle, n, dim, nlist })
}

/// src/lib.rs:2913 `build_ivf_index_gpu`; `devices` replaces `R::Device` (one ordinal, or several: lists sharded over the box).
pub fn build_ivf_index_gpu(mat: MatRef<f32>, nlist: Option<usize>, k_means_params: Option<KMeansTrainingParams>, dist_metric: &str, seed: usize,
                           verbose: bool, devices: &[i32]) -> Result<IvfIndexB200, AnnSearchErrors> {
    build_ivf(mat, nlist, k_means_params, dist_metric, seed, verbose, ANNB_F32, devices)
}
/// src/lib.rs:2100 `build_ivf_bf16_index`.
pub fn build_ivf_bf16_index(mat: MatRef<f32>, nlist: Option<usize>, k_means_params: Option<KMeansTrainingParams>, dist_metric: &str, seed: usize,
                            verbose: bool, devices: &[i32]) -> Result<IvfIndexB200, AnnSearchErrors> {
    build_ivf(mat, nlist, k_means_params, dist_metric, seed, verbose, ANNB_BF16, devices)
}
/// src/lib.rs:2196 `build_ivf_sq8_index`.
pub fn build_ivf_sq8_index(mat: MatRef<f32>, nlist: Option<usize>, k_means_params: Option<KMeansTrainingParams>, dist_metric: &str, seed: usize,
                           verbose: bool, devices: &[i32]) -> Result<IvfIndexB200, AnnSearchErrors> {
    build_ivf(mat, nlist, k_means_params, dist_metric, seed, verbose, ANNB_SQ8, devices)
}
/// src/lib.rs:2142 / 2238: the BF16 / SQ8 query functions share the f32 entry point (queries are f32 at the ABI).
pub use query_ivf_index_gpu as query_ivf_bf16_index;
pub use query_ivf_index_gpu as query_ivf_sq8_index;
pub use query_ivf_index_gpu_self as query_ivf_bf16_self;
pub use query_ivf_index_gpu_self as query_ivf_sq8_self;

/// `build_exhaustive_index_gpu` over several GPUs of the box: rows sharded behind one handle (annb_flat_create_multi);
/// `query_exhaustive_index_gpu[_self]` take the result unchanged.
pub fn build_exhaustive_index_multi_gpu(mat: MatRef<f32>, dist_metric: &str, dtype: c_int, devices: &[i32]) -> Result<ExhaustiveIndexB200, AnnSearchErrors> {
    let (flat, n, dim) = matrix_to_flat(mat);
    let mut handle = std::ptr::null_mut();
    check(unsafe { annb_flat_create_multi(&mut handle, flat.as_ptr(), n as u64, dim as u32, dtype, metric_or_default(dist_metric),
                                          devices.as_ptr(), devices.len() as c_int) })?;
    Ok(ExhaustiveIndexB200 { handle, n, dim })
}

/// `KnnGraphGpu<T>` (src/gpu/nndescent_gpu.rs:2418-2446) filled by the exhaustive self search: the hand-off struct of
/// `build_nsg_from_gpu_knn` (src/lib.rs:3330) and the raw-kNN consumers.  Same fields, same sentinel convention.
pub struct KnnGraphGpu {
    pub vectors_flat: Vec<f32>,
    pub dim: usize,
    pub n: usize,
    pub k: usize,
    pub norms: Vec<f32>,
    pub metric: c_int,
    pub knn_graph: Vec<(usize, f32)>,
    pub converged: bool,
}

/// src/lib.rs:3201 `build_knn_graph_gpu`, exact instead of NN-Descent (the NN-Descent tuning arguments have no meaning here).
pub fn build_knn_graph_gpu(mat: MatRef<f32>, dist_metric: &str, k: Option<usize>, devices: &[i32]) -> Result<KnnGraphGpu, AnnSearchErrors> {
    let k = k.unwrap_or(30);
    let metric = metric_or_default(dist_metric);
    let index = if devices.len() > 1 { build_exhaustive_index_multi_gpu(mat, dist_metric, ANNB_F32, devices)? }
                else { build_exhaustive_index_gpu(mat, dist_metric, devices.first().copied().unwrap_or(0))? };
    let (flat, n, dim) = matrix_to_flat(mat);
    let (mut pid, mut dist) = (vec![0u64; n * k], vec![0f32; n * k]);
    check(unsafe { annb_flat_knn_graph(index.handle, 0, n as u64, k as u32, pid.as_mut_ptr(), dist.as_mut_ptr(), std::ptr::null_mut()) })?;
    let norms = if metric == ANNB_COSINE { flat.chunks_exact(dim).map(l2_norm).collect() } else { Vec::new() };
    let knn_graph = pid.iter().zip(&dist).map(|(&p, &d)| (p as usize, d)).collect();
    Ok(KnnGraphGpu { vectors_flat: flat, dim, n, k, norms, metric, knn_graph, converged: true })
}
