// annb200.hpp -- header-only C++17 host mirror of the ann-search-rs free-function API for the flat and IVF
// families, on top of the C ABI in include/annb200.h.  (The reference is a Rust crate; no Rust toolchain exists in the
// build image, so the compiled-language host side above the C ABI is C++.  The Rust shim a maintainer would add is in
// ann-search-rs_b200/rust/ and INTEGRATION.md.)
//
// Names, argument order and meaning follow src/lib.rs of the reference:
//   build_exhaustive_index_gpu / query_exhaustive_index_gpu / query_exhaustive_index_gpu_self      lib.rs:2813-2911
//   build_ivf_index_gpu / query_ivf_index_gpu / query_ivf_index_gpu_self                          lib.rs:2913-3002
//   build_exhaustive_{bf16,sq8}_index / query_exhaustive_{bf16,sq8}_{index,self}                   lib.rs:1702-1871
//   query_ivf_{bf16,sq8}_{index,self}                                                              lib.rs:2142-2290
// faer::MatRef<T>  -> annb200::MatRef (pointer + strides, rows = samples)
// KnnOptionResult  -> annb200::KnnResult {indices, optional distances}; errors -> annb200::AnnSearchError (variant()).
#pragma once
#include <cstdint>
#include <cstdio>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/annb200.h"

namespace annb200 {

enum class ErrorVariant { DimensionMismatch, DistanceNotSupported, TooFewSamplesForCentroids, InvalidArgument, Cuda, Nccl, OutOfMemory, Unsupported, Unknown };

// Mirror of AnnSearchErrors (src/errors.rs).
class AnnSearchError : public std::runtime_error {
  public:
    AnnSearchError(int code, const std::string& msg) : std::runtime_error(msg), code_(code) {}
    int code() const { return code_; }
    ErrorVariant variant() const {
        switch (code_) {
            case ANNB_ERR_DIMENSION_MISMATCH: return ErrorVariant::DimensionMismatch;
            case ANNB_ERR_DISTANCE_NOT_SUPPORTED: return ErrorVariant::DistanceNotSupported;
            case ANNB_ERR_TOO_FEW_SAMPLES: return ErrorVariant::TooFewSamplesForCentroids;
            case ANNB_ERR_INVALID_ARGUMENT: return ErrorVariant::InvalidArgument;
            case ANNB_ERR_CUDA: return ErrorVariant::Cuda;
            case ANNB_ERR_NCCL: return ErrorVariant::Nccl;
            case ANNB_ERR_OUT_OF_MEMORY: return ErrorVariant::OutOfMemory;
            case ANNB_ERR_UNSUPPORTED: return ErrorVariant::Unsupported;
            default: return ErrorVariant::Unknown;
        }
    }

  private:
    int code_;
};

inline void check(int status) {
    if (status != ANNB_OK) throw AnnSearchError(status, annb_last_error());
}

// faer::MatRef<f32>: arbitrary strides, column-major by default in faer.
struct MatRef {
    const float* ptr = nullptr;
    size_t nrows = 0, ncols = 0;
    ptrdiff_t row_stride = 0, col_stride = 0;
    static MatRef row_major(const float* p, size_t r, size_t c) { return {p, r, c, static_cast<ptrdiff_t>(c), 1}; }
    static MatRef col_major(const float* p, size_t r, size_t c) { return {p, r, c, 1, static_cast<ptrdiff_t>(r)}; }
    float at(size_t i, size_t j) const { return ptr[static_cast<ptrdiff_t>(i) * row_stride + static_cast<ptrdiff_t>(j) * col_stride]; }
};

// matrix_to_flat (src/utils/mod.rs:44-68): owned row-major copy.
inline std::vector<float> matrix_to_flat(const MatRef& m) {
    std::vector<float> out(m.nrows * m.ncols);
    if (m.col_stride == 1 && m.row_stride == static_cast<ptrdiff_t>(m.ncols)) {
        std::copy(m.ptr, m.ptr + out.size(), out.begin());
    } else {
        for (size_t i = 0; i < m.nrows; i++)
            for (size_t j = 0; j < m.ncols; j++) out[i * m.ncols + j] = m.at(i, j);
    }
    return out;
}

// parse_ann_dist + the fallback policy of the free functions (src/utils/dist.rs:63-70, src/lib.rs:274-277).
inline int metric_or_default(const std::string& dist_metric) {
    int m = annb_parse_metric(dist_metric.c_str());
    if (m < 0) {
        std::fprintf(stderr, "  Unknown distance metric '%s', defaulting to Euclidean\n", dist_metric.c_str());
        return ANNB_L2;
    }
    return m;
}

struct KnnResult {
    std::vector<std::vector<size_t>> indices;
    std::optional<std::vector<std::vector<float>>> distances;
};

namespace detail {
inline KnnResult unpack(const std::vector<uint64_t>& ids, const std::vector<float>& dist, const std::vector<uint32_t>& cnt, size_t nq, size_t k,
                        bool return_dist) {
    KnnResult r;
    r.indices.resize(nq);
    if (return_dist) r.distances.emplace(nq);
    for (size_t i = 0; i < nq; i++) {
        r.indices[i].assign(ids.begin() + i * k, ids.begin() + i * k + cnt[i]);
        if (return_dist) (*r.distances)[i].assign(dist.begin() + i * k, dist.begin() + i * k + cnt[i]);
    }
    return r;
}
}  // namespace detail

class IndexHandle {
  public:
    IndexHandle() = default;
    explicit IndexHandle(annb_index* h) : h_(h) {}
    IndexHandle(IndexHandle&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    IndexHandle& operator=(IndexHandle&& o) noexcept {
        if (this != &o) { reset(); h_ = o.h_; o.h_ = nullptr; }
        return *this;
    }
    IndexHandle(const IndexHandle&) = delete;
    IndexHandle& operator=(const IndexHandle&) = delete;
    ~IndexHandle() { reset(); }
    annb_index* get() const { return h_; }
    annb_index_info info() const { annb_index_info i{}; check(annb_index_get_info(h_, &i)); return i; }
    // (ram, vram) as IvfIndexGpu::memory_usage_bytes (src/gpu/ivf_gpu.rs:590-604)
    std::pair<size_t, size_t> memory_usage_bytes() const { auto i = info(); return {i.host_bytes, i.device_bytes}; }

  private:
    void reset() { if (h_) annb_destroy(h_); h_ = nullptr; }
    annb_index* h_ = nullptr;
};

// ExhaustiveIndexGpu (src/gpu/exhaustive_gpu.rs:17-33) and its BF16 / SQ8 twins.
class ExhaustiveIndexB200 : public IndexHandle {
  public:
    using IndexHandle::IndexHandle;
    static ExhaustiveIndexB200 create(const MatRef& data, int metric, int dtype, int device = 0) {
        auto flat = matrix_to_flat(data);
        annb_index* h = nullptr;
        check(annb_flat_create(&h, flat.data(), data.nrows, static_cast<uint32_t>(data.ncols), dtype, metric, nullptr, 0, device));
        return ExhaustiveIndexB200(h);
    }
    KnnResult query_batch(const MatRef& q, size_t k, bool return_dist) const {
        auto flat = matrix_to_flat(q);
        std::vector<uint64_t> ids(q.nrows * k);
        std::vector<float> dist(return_dist ? q.nrows * k : 0);
        std::vector<uint32_t> cnt(q.nrows);
        check(annb_flat_search(get(), flat.data(), q.nrows, static_cast<uint32_t>(q.ncols), static_cast<uint32_t>(k), ids.data(),
                               return_dist ? dist.data() : nullptr, cnt.data()));
        return detail::unpack(ids, dist, cnt, q.nrows, k, return_dist);
    }
    KnnResult generate_knn(size_t k, bool return_dist) const {
        const size_t n = info().n;
        std::vector<uint64_t> ids(n * k);
        std::vector<float> dist(return_dist ? n * k : 0);
        std::vector<uint32_t> cnt(n);
        check(annb_flat_search_self(get(), 0, n, static_cast<uint32_t>(k), ids.data(), return_dist ? dist.data() : nullptr, cnt.data()));
        return detail::unpack(ids, dist, cnt, n, k, return_dist);
    }
};

// IvfIndexGpu (src/gpu/ivf_gpu.rs:153-181) and its BF16 / SQ8 twins.  Built from the contents of the reference's index
// struct (annb_ivf_create); the host-side build steps live with the caller (see python/annb200 for a worked mirror).
class IvfIndexB200 : public IndexHandle {
  public:
    using IndexHandle::IndexHandle;
    KnnResult query_batch(const MatRef& q, size_t k, std::optional<size_t> nprobe, bool return_dist) const {
        auto flat = matrix_to_flat(q);
        std::vector<uint64_t> ids(q.nrows * k);
        std::vector<float> dist(return_dist ? q.nrows * k : 0);
        std::vector<uint32_t> cnt(q.nrows);
        check(annb_ivf_search(get(), flat.data(), q.nrows, static_cast<uint32_t>(q.ncols), static_cast<uint32_t>(k),
                              static_cast<uint32_t>(nprobe.value_or(0)), ids.data(), return_dist ? dist.data() : nullptr, cnt.data()));
        return detail::unpack(ids, dist, cnt, q.nrows, k, return_dist);
    }
    KnnResult generate_knn(size_t k, std::optional<size_t> nprobe, bool return_dist) const {
        const size_t n = info().n;
        std::vector<uint64_t> ids(n * k, UINT64_MAX);
        std::vector<float> dist(return_dist ? n * k : 0);
        std::vector<uint32_t> cnt(n, 0);
        check(annb_ivf_search_self(get(), 0, n, static_cast<uint32_t>(k), static_cast<uint32_t>(nprobe.value_or(0)), 1, ids.data(),
                                   return_dist ? dist.data() : nullptr, cnt.data()));
        return detail::unpack(ids, dist, cnt, n, k, return_dist);
    }
    // KnnValidation::validate_index (src/utils/mod.rs:210-242): recall@k against an exhaustive search over the index's own vectors
    // on the stored vectors at `positions` (drawn by the caller, as the reference draws them with its StdRng).
    double validate_index(size_t k, const std::vector<uint64_t>& positions) const {
        double recall = 0.0;
        check(annb_ivf_validate(get(), positions.data(), positions.size(), static_cast<uint32_t>(k), 0, &recall));
        return recall;
    }
};

// ---- free functions, src/lib.rs ---------------------------------------------------------------------------------------
inline ExhaustiveIndexB200 build_exhaustive_index_gpu(const MatRef& mat, const std::string& dist_metric, int device = 0) {
    return ExhaustiveIndexB200::create(mat, metric_or_default(dist_metric), ANNB_F32, device);
}
inline ExhaustiveIndexB200 build_exhaustive_bf16_index(const MatRef& mat, const std::string& dist_metric, int device = 0) {
    return ExhaustiveIndexB200::create(mat, metric_or_default(dist_metric), ANNB_BF16, device);
}
inline ExhaustiveIndexB200 build_exhaustive_sq8_index(const MatRef& mat, const std::string& dist_metric, int device = 0) {
    return ExhaustiveIndexB200::create(mat, metric_or_default(dist_metric), ANNB_SQ8, device);
}
inline KnnResult query_exhaustive_index_gpu(const MatRef& query_mat, const ExhaustiveIndexB200& index, size_t k, bool return_dist, bool /*verbose*/ = false) {
    return index.query_batch(query_mat, k, return_dist);
}
inline KnnResult query_exhaustive_index_gpu_self(const ExhaustiveIndexB200& index, size_t k, bool return_dist, bool /*verbose*/ = false) {
    return index.generate_knn(k, return_dist);
}
inline KnnResult query_ivf_index_gpu(const MatRef& query_mat, const IvfIndexB200& index, size_t k, std::optional<size_t> nprobe,
                                     std::optional<size_t> /*nquery: batching is internal*/, bool return_dist, bool /*verbose*/ = false) {
    return index.query_batch(query_mat, k, nprobe, return_dist);
}
inline KnnResult query_ivf_index_gpu_self(const IvfIndexB200& index, size_t k, std::optional<size_t> nprobe, std::optional<size_t> /*nquery*/,
                                          bool return_dist, bool /*verbose*/ = false) {
    return index.generate_knn(k, nprobe, return_dist);
}

// Lloyd loop of train_centroids on the device (src/utils/k_means_utils.rs:1572-1700, unbalanced parallel_lloyd) from the
// caller's initial centroids [n_centroids * dim]; returns the trained centroids, *iters (optional) = updates performed.
inline std::vector<float> kmeans_lloyd(const MatRef& train, std::vector<float> centroids, size_t n_centroids, const std::string& dist_metric,
                                       size_t max_iters = 30, uint32_t* iters = nullptr, int device = 0) {
    const std::vector<float> flat = matrix_to_flat(train);
    if (centroids.size() != n_centroids * train.ncols) throw AnnSearchError(ANNB_ERR_DIMENSION_MISMATCH, "centroids must be n_centroids x dim");
    check(annb_kmeans_lloyd(flat.data(), train.nrows, static_cast<uint32_t>(train.ncols), centroids.data(), static_cast<uint32_t>(n_centroids),
                            metric_or_default(dist_metric), static_cast<uint32_t>(max_iters), iters, device));
    return centroids;
}

}  // namespace annb200
