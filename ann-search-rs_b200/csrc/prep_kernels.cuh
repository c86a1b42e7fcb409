// prep_kernels.cuh -- index-construction and query-preparation kernels.
// Each mirrors one step of the reference constructors with the same arithmetic, so the
// device-built index holds exactly the bytes the CPU index would hold.
#pragma once
#include "common.cuh"
#include "refdist.cuh"

namespace annb {

// [n][src_row_bytes] -> [n][dst_row_bytes], zero padded (dst_row_bytes % 16 == 0).
__global__ void pad_rows_kernel(const uint8_t* __restrict__ src, uint32_t src_row_bytes, uint8_t* __restrict__ dst,
                                uint32_t dst_row_bytes, uint64_t n) {
    const uint64_t total = n * dst_row_bytes;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        uint64_t r = i / dst_row_bytes;
        uint32_t c = static_cast<uint32_t>(i - r * dst_row_bytes);
        dst[i] = (c < src_row_bytes) ? src[r * src_row_bytes + c] : 0;
    }
}

// calculate_l2_norm per row (src/utils/dist.rs:2339-2360), rows padded f32 with pitch ld floats.
__global__ void row_norms_f32_kernel(const float* __restrict__ rows, uint32_t ld, uint32_t dim, uint64_t n, float* __restrict__ out,
                                     int squared) {
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    float s = ref_dot_self_f32(rows + i * ld, dim);
    out[i] = squared ? s : __fsqrt_rn(s);
}

// normalise_vector per row (src/utils/dist.rs:5336-5344).
__global__ void normalise_rows_f32_kernel(float* __restrict__ rows, uint32_t ld, uint32_t dim, uint64_t n) {
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    float* v = rows + i * ld;
    float nrm = __fsqrt_rn(ref_dot_self_f32(v, dim));
    if (nrm > 0.0f)
        for (uint32_t e = 0; e < dim; e++) v[e] = __fdiv_rn(v[e], nrm);
}

// encode_bf16_quantisation (src/quantised/quantisers.rs:31-38): RNE.  f32 pitch ld_src floats ->
// bf16 pitch ld_dst elements, zero padded.
__global__ void encode_bf16_kernel(const float* __restrict__ src, uint32_t ld_src, uint32_t dim, uint16_t* __restrict__ dst,
                                   uint32_t ld_dst, uint64_t n) {
    const uint64_t total = n * ld_dst;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        uint64_t r = i / ld_dst;
        uint32_t c = static_cast<uint32_t>(i - r * ld_dst);
        uint16_t v = 0;
        if (c < dim) {
            uint32_t x = __float_as_uint(src[r * ld_src + c]);
            if ((x & 0x7FFFFFFFu) > 0x7F800000u) v = static_cast<uint16_t>((x >> 16) | 0x0040u);
            else if ((x & 0x8000u) != 0 && (x & 0x17FFFu) != 0) v = static_cast<uint16_t>((x >> 16) + 1);
            else v = static_cast<uint16_t>(x >> 16);
        }
        dst[i] = v;
    }
}

// bf16 -> f32 widening (decode_bf16_quantisation, quantisers.rs:50-62); same pitch in elements.
__global__ void decode_bf16_kernel(const uint16_t* __restrict__ src, uint32_t ld_src, float* __restrict__ dst, uint32_t ld_dst,
                                   uint32_t dim, uint64_t n) {
    const uint64_t total = n * ld_dst;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        uint64_t r = i / ld_dst;
        uint32_t c = static_cast<uint32_t>(i - r * ld_dst);
        dst[i] = (c < dim) ? bf16_bits_to_f32(src[r * ld_src + c]) : 0.0f;
    }
}

// ScalarQuantiser::train (src/quantised/quantisers.rs:123-146): per-dimension max |x|, then /128.
__global__ void sq8_absmax_kernel(const float* __restrict__ rows, uint32_t ld, uint32_t dim, uint64_t n, uint32_t* __restrict__ maxbits) {
    // one thread column per dimension inside a block row-stripe; non-negative floats order like their bits
    const uint32_t d = blockIdx.y * blockDim.x + threadIdx.x;
    if (d >= dim) return;
    uint32_t m = 0;
    for (uint64_t r = blockIdx.x; r < n; r += gridDim.x) {
        float a = fabsf(rows[r * ld + d]);
        if (!(a != a)) m = max(m, __float_as_uint(a));  // Float::max ignores NaN
    }
    atomicMax(maxbits + d, m);
}
__global__ void sq8_scales_kernel(const uint32_t* __restrict__ maxbits, uint32_t dim, float* __restrict__ scales) {
    uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= dim) return;
    float mx = __uint_as_float(maxbits[d]);
    scales[d] = (mx <= 0.0f) ? 1.0f : __fdiv_rn(mx, 128.0f);
}

__device__ __forceinline__ int8_t sq8_encode_one(float val, float scale) {
    // ScalarQuantiser::encode (src/quantised/quantisers.rs:148-165)
    float scaled = __fdiv_rn(val, scale);
    if (scaled != scaled) return 0;  // to_i8() of NaN -> None -> 0
    float sg = (__float_as_uint(scaled) >> 31) ? -1.0f : 1.0f;  // f32::signum (+0 -> 1, -0 -> -1)
    float rounded = __fadd_rn(scaled, __fmul_rn(0.5f, sg));
    float clamped = fmaxf(fminf(rounded, 127.0f), -128.0f);
    return static_cast<int8_t>(static_cast<int>(clamped));  // truncation toward zero
}
__global__ void sq8_encode_kernel(const float* __restrict__ src, uint32_t ld_src, uint32_t dim, const float* __restrict__ scales,
                                  int8_t* __restrict__ dst, uint32_t ld_dst, uint64_t n) {
    const uint64_t total = n * ld_dst;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        uint64_t r = i / ld_dst;
        uint32_t c = static_cast<uint32_t>(i - r * ld_dst);
        dst[i] = (c < dim) ? sq8_encode_one(src[r * ld_src + c], scales[c]) : 0;
    }
}
// ScalarQuantiser::decode (quantisers.rs:167-175): code * scale.
__global__ void sq8_decode_kernel(const int8_t* __restrict__ src, uint32_t ld_src, const float* __restrict__ scales,
                                  float* __restrict__ dst, uint32_t ld_dst, uint32_t dim, uint64_t n) {
    const uint64_t total = n * ld_dst;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        uint64_t r = i / ld_dst;
        uint32_t c = static_cast<uint32_t>(i - r * ld_dst);
        dst[i] = (c < dim) ? __fmul_rn(__int2float_rn(src[r * ld_src + c]), scales[c]) : 0.0f;
    }
}
// sum of squared codes per row (exhaustive_sq8.rs:132-141).
__global__ void sq8_row_norms_kernel(const int8_t* __restrict__ rows, uint32_t ld, uint32_t dim, uint64_t n, int32_t* __restrict__ out) {
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    int32_t s = 0;
    for (uint32_t e = 0; e < dim; e++) {
        int32_t v = rows[i * ld + e];
        s += v * v;
    }
    out[i] = s;
}

// Row gather / result scatter used by the exact fallback of the tensor paths.
__global__ void gather_rows_kernel(const uint8_t* __restrict__ src, uint32_t row_bytes, const uint32_t* __restrict__ list, uint32_t count,
                                   uint8_t* __restrict__ dst) {
    const uint32_t cpr = row_bytes >> 4;
    const uint64_t total = static_cast<uint64_t>(count) * cpr;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint32_t r = static_cast<uint32_t>(i / cpr), c = static_cast<uint32_t>(i - static_cast<uint64_t>(r) * cpr);
        reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src + static_cast<uint64_t>(list[r]) * row_bytes)[c];
    }
}
// dst[i][:] = src[list[i]][:] for rows of `ld` 32-bit words (probe lists of the queries sent to the exact fallback).
__global__ void gather_u32_rows_kernel(const uint32_t* __restrict__ src, uint32_t ld, const uint32_t* __restrict__ list, uint32_t n, uint32_t* __restrict__ dst) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<uint64_t>(n) * ld) return;
    const uint64_t r = i / ld;
    dst[i] = src[static_cast<uint64_t>(list[r]) * ld + (i - r * ld)];
}
// Probe lists [nq][src_pitch] -> caller layout [nq][dst_pitch]; flags a query whose list does not fit.
__global__ void export_probes_kernel(const uint32_t* __restrict__ probes, uint32_t src_pitch, const uint32_t* __restrict__ n_probes, uint64_t nq,
                                     uint32_t* __restrict__ out_probes, uint32_t dst_pitch, uint32_t* __restrict__ out_n, uint32_t* __restrict__ overflow) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= nq * dst_pitch) return;
    const uint64_t q = i / dst_pitch;
    const uint32_t r = static_cast<uint32_t>(i - q * dst_pitch);
    const uint32_t np = n_probes[q];
    out_probes[i] = (r < np && r < src_pitch) ? probes[q * src_pitch + r] : 0xFFFFFFFFu;
    if (r == 0) {
        out_n[q] = min(np, dst_pitch);
        if (np > dst_pitch) atomicExch(overflow, 1u);
    }
}
__global__ void scatter_results_kernel(const uint32_t* __restrict__ list, uint32_t count, uint32_t k, const uint64_t* __restrict__ ids,
                                       const float* __restrict__ dist, const uint32_t* __restrict__ cnt, uint64_t* __restrict__ out_ids,
                                       float* __restrict__ out_dist, uint32_t* __restrict__ out_cnt) {
    const uint64_t total = static_cast<uint64_t>(count) * k;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint32_t r = static_cast<uint32_t>(i / k), j = static_cast<uint32_t>(i - static_cast<uint64_t>(r) * k);
        const uint64_t o = static_cast<uint64_t>(list[r]) * k + j;
        out_ids[o] = ids[i];
        if (out_dist) out_dist[o] = dist[i];
        if (out_cnt && j == 0) out_cnt[list[r]] = cnt[r];
    }
}

// direct_assign shortcut norms (src/utils/k_means_utils.rs:2134-2156): |c|^2 via dot_simd (L2) or
// 1/norm (cosine; 0 when the norm is 0) where norm is the caller-supplied centroid norm, or when none
// is supplied the sequential-fold norm IvfIndex::build computes (src/cpu/ivf.rs:193-206).
__global__ void assign_aux_kernel(const float* __restrict__ cent, uint32_t ld, uint32_t dim, uint32_t nlist, int cosine,
                                  const float* __restrict__ cnorms, float* __restrict__ aux) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nlist) return;
    const float* v = cent + static_cast<uint64_t>(c) * ld;
    if (!cosine) aux[c] = ref_dot_self_f32(v, dim);
    else {
        float nrm = cnorms ? cnorms[c] : seq_norm<4>(reinterpret_cast<const uint8_t*>(v), dim);
        aux[c] = (nrm > 0.0f) ? __fdiv_rn(1.0f, nrm) : 0.0f;
    }
}
// ---- Lloyd iteration pieces (parallel_lloyd, src/utils/k_means_utils.rs:1572-1700) ----
// changed = #{i : assign[i] != prev[i]}; prev <- assign
__global__ void kmeans_changed_kernel(const uint32_t* __restrict__ assign, uint32_t* __restrict__ prev, uint64_t n, unsigned long long* __restrict__ changed) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    const bool diff = i < n && assign[i] != prev[i];
    if (i < n) prev[i] = assign[i];
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, diff);
    if ((threadIdx.x & 31u) == 0 && m) atomicAdd(changed, static_cast<unsigned long long>(__popc(m)));
}
// per-cluster sums (f64 accumulators: the order of the reference's per-thread f32 partial sums depends on the rayon
// pool size and is not reproducible; f64 makes the mean independent of the accumulation order to ~1e-16) and counts
__global__ void kmeans_accumulate_kernel(const float* __restrict__ x, uint32_t ld, uint32_t dim, uint64_t n, const uint32_t* __restrict__ assign,
                                         double* __restrict__ sums, uint32_t* __restrict__ counts) {
    const uint64_t total = n * dim;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t r = i / dim;
        const uint32_t e = static_cast<uint32_t>(i - r * dim);
        const uint32_t c = assign[r];
        atomicAdd(sums + static_cast<uint64_t>(c) * dim + e, static_cast<double>(x[r * ld + e]));
        if (e == 0) atomicAdd(counts + c, 1u);
    }
}
// centroid <- sum / count for non-empty clusters; empty clusters keep their centroid (k_means_utils.rs:1657-1666)
__global__ void kmeans_update_kernel(const double* __restrict__ sums, const uint32_t* __restrict__ counts, float* __restrict__ cent, uint32_t ld, uint32_t dim,
                                     uint32_t k) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<uint64_t>(k) * dim) return;
    const uint32_t c = static_cast<uint32_t>(i / dim), e = static_cast<uint32_t>(i - static_cast<uint64_t>(c) * dim);
    if (counts[c] > 0) cent[static_cast<uint64_t>(c) * ld + e] = static_cast<float>(sums[i] / static_cast<double>(counts[c]));
}

// Sequential-fold norms of f32 rows (centroid norms, src/cpu/ivf.rs:193-206).
__global__ void seq_norms_kernel(const float* __restrict__ rows, uint32_t ld, uint32_t dim, uint64_t n, float* __restrict__ out) {
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    out[i] = seq_norm<4>(reinterpret_cast<const uint8_t*>(rows + i * ld), dim);
}

// kNN-graph rows from self-search rows (KnnGraphGpu, src/gpu/nndescent_gpu.rs:2418-2446, 2611-2667): row i of the
// [nq][k + 1] search result loses its own id (wherever it sits -- exact duplicates may precede it) and keeps the first k
// survivors, ascending by distance; unfilled slots carry the reference's sentinel pair (SENTINEL_PID = u32::MAX >> 1, T::MAX).
__global__ void knn_graph_rows_kernel(const uint64_t* __restrict__ ids, const float* __restrict__ dist, uint64_t nq, uint32_t k1, uint64_t self0,
                                      uint64_t* __restrict__ out_ids, float* __restrict__ out_dist, uint32_t* __restrict__ out_cnt) {
    const uint64_t q = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (q >= nq) return;
    const uint32_t k = k1 - 1;
    const uint64_t self = self0 + q;
    uint32_t w = 0;
    for (uint32_t j = 0; j < k1 && w < k; j++) {
        const uint64_t id = ids[q * k1 + j];
        if (id == 0xFFFFFFFFFFFFFFFFull || id == self) continue;
        out_ids[q * k + w] = id;
        out_dist[q * k + w] = dist[q * k1 + j];
        w++;
    }
    if (out_cnt) out_cnt[q] = w;
    for (; w < k; w++) {
        out_ids[q * k + w] = 0x7FFFFFFFull;
        out_dist[q * k + w] = 3.402823466e+38f;
    }
}

}  // namespace annb
