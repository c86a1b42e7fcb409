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
