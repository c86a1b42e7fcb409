// refdist.cuh -- distances computed in the *reference's floating-point order*.
//
// The CPU reference accumulates in 8 SIMD lanes over consecutive 8-element
// chunks, reduces the 8 lanes with a fixed horizontal-add tree and then folds
// the (dim % 8) tail sequentially:
//   f32 : src/utils/dist.rs:306-330 (euclid), 587-609 (dot)      mul + add, wide::f32x8::reduce_add
//   bf16: src/utils/dist.rs:3392-3416, 3615-3636, 4118-4150, 4322-4357   fused multiply-add, hsum_f32_avx2
//   sq8 : src/utils/dist.rs:5015-5077                             exact i32 arithmetic
// One GPU thread owns one (row, query) pair and walks the row in that exact
// order with IEEE round-to-nearest intrinsics (no contraction), so the f32 it
// produces is bit-identical to the oracle's.  These functions serve the exact
// SIMT kernels and the re-rank step of the tensor-core path.
#pragma once
#include "common.cuh"

namespace annb {

enum { MET_L2 = 0, MET_COS = 1, MET_COS_PRENORM = 2, MET_DOT = 3 };
enum { QT_F32 = 0, QT_BF16 = 1, QT_I8 = 2 };

__device__ __forceinline__ float bf16_bits_to_f32(uint32_t h) { return __uint_as_float(h << 16); }

// 8 consecutive elements of a row as f32 (exact widening for bf16).
template <int ELEM>  // 4 = f32, 2 = bf16
__device__ __forceinline__ void load8(const uint8_t* p, float v[8]) {
    if (ELEM == 4) {
        float4 a = *reinterpret_cast<const float4*>(p);
        float4 b = *reinterpret_cast<const float4*>(p + 16);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        uint4 a = *reinterpret_cast<const uint4*>(p);
        v[0] = bf16_bits_to_f32(a.x & 0xFFFFu); v[1] = bf16_bits_to_f32(a.x >> 16);
        v[2] = bf16_bits_to_f32(a.y & 0xFFFFu); v[3] = bf16_bits_to_f32(a.y >> 16);
        v[4] = bf16_bits_to_f32(a.z & 0xFFFFu); v[5] = bf16_bits_to_f32(a.z >> 16);
        v[6] = bf16_bits_to_f32(a.w & 0xFFFFu); v[7] = bf16_bits_to_f32(a.w >> 16);
    }
}
template <int ELEM>
__device__ __forceinline__ float load1(const uint8_t* row, int e) {
    if (ELEM == 4) return reinterpret_cast<const float*>(row)[e];
    return bf16_bits_to_f32(reinterpret_cast<const uint16_t*>(row)[e]);
}

// wide 1.4.0 f32x8::reduce_add (AVX): (a0+a4, a1+a5, a2+a6, a3+a7) -> (s0+s2, s1+s3) -> t0+t1
__device__ __forceinline__ float hsum_wide(const float a[8]) {
    float s0 = __fadd_rn(a[0], a[4]), s1 = __fadd_rn(a[1], a[5]);
    float s2 = __fadd_rn(a[2], a[6]), s3 = __fadd_rn(a[3], a[7]);
    return __fadd_rn(__fadd_rn(s0, s2), __fadd_rn(s1, s3));
}
// hsum_f32_avx2 -> hsum_f32_sse (src/utils/dist.rs:3167-3178, 3211-3220): (s0+s1) + (s2+s3)
__device__ __forceinline__ float hsum_bf16path(const float a[8]) {
    float s0 = __fadd_rn(a[0], a[4]), s1 = __fadd_rn(a[1], a[5]);
    float s2 = __fadd_rn(a[2], a[6]), s3 = __fadd_rn(a[3], a[7]);
    return __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3));
}

// Raw accumulation (squared distance for MET_L2, dot product otherwise) of one
// stored row against NQ queries held in shared memory.
//   RELEM : bytes per row element (4 f32, 2 bf16);  QELEM : bytes per query element.
//   FMA   : bf16 kernels of the reference use _mm256_fmadd_ps, f32 kernels do not.
//   DEEP  : four chunks per loop trip with their row loads issued first, so that a latency-bound caller (one lane walking a row that sits in L2) keeps
//           eight 128-bit row loads in flight instead of the compiler's two; the order of the arithmetic is unchanged.
template <int RELEM, int QELEM, bool L2, int NQ, bool DEEP = false>
__device__ __forceinline__ void accumulate_fp(const uint8_t* __restrict__ row, const uint8_t* __restrict__ q,
                                              uint32_t q_stride, int dim, float out[NQ]) {
    constexpr bool FMA = (RELEM == 2);
    float acc[NQ][8];
#pragma unroll
    for (int i = 0; i < NQ; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.0f;
    const int chunks = dim >> 3;
    auto chunk_x = [&](int c, const float x[8]) {
#pragma unroll
        for (int i = 0; i < NQ; i++) {
            float y[8];
            load8<QELEM>(q + i * q_stride + c * 8 * QELEM, y);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (L2) {
                    float d = __fsub_rn(x[j], y[j]);
                    acc[i][j] = FMA ? __fmaf_rn(d, d, acc[i][j]) : __fadd_rn(acc[i][j], __fmul_rn(d, d));
                } else {
                    acc[i][j] = FMA ? __fmaf_rn(x[j], y[j], acc[i][j]) : __fadd_rn(acc[i][j], __fmul_rn(x[j], y[j]));
                }
            }
        }
    };
    auto chunk = [&](int c) {
        float x[8];
        load8<RELEM>(row + c * 8 * RELEM, x);
        chunk_x(c, x);
    };
    if constexpr (DEEP) {
        int c = 0;
        for (; c + 4 <= chunks; c += 4) {
            float x[4][8];
#pragma unroll
            for (int u = 0; u < 4; u++) load8<RELEM>(row + (c + u) * 8 * RELEM, x[u]);   // the four chunks' row loads go out together ...
            asm volatile("" ::: "memory");                                              // ... (compiler barrier: they are not sunk into the arithmetic)
#pragma unroll
            for (int u = 0; u < 4; u++) chunk_x(c + u, x[u]);
        }
        for (; c < chunks; c++) chunk(c);
    } else {
        for (int c = 0; c < chunks; c++) chunk(c);
    }
#pragma unroll
    for (int i = 0; i < NQ; i++) {
        float sum = FMA ? hsum_bf16path(acc[i]) : hsum_wide(acc[i]);
        for (int e = chunks * 8; e < dim; e++) {  // scalar tail: `sum += d * d` (not fused)
            float xv = load1<RELEM>(row, e);
            float yv = load1<QELEM>(q + i * q_stride, e);
            if (L2) {
                float d = __fsub_rn(xv, yv);
                sum = __fadd_rn(sum, __fmul_rn(d, d));
            } else {
                sum = __fadd_rn(sum, __fmul_rn(xv, yv));
            }
        }
        out[i] = sum;
    }
}

// SQ8 code space: exact integers.  Rows are zero padded to 16 codes, so whole
// 16-byte chunks can go through dp4a.  Returns dot(q, x) per query and x.x.
template <int NQ>
__device__ __forceinline__ void accumulate_i8(const uint8_t* __restrict__ row, const uint8_t* __restrict__ q,
                                              uint32_t q_stride, int dim, int32_t dot[NQ], int32_t& xx) {
#pragma unroll
    for (int i = 0; i < NQ; i++) dot[i] = 0;
    xx = 0;
    const int chunks = (dim + 15) >> 4;
    for (int c = 0; c < chunks; c++) {
        int4 x = *reinterpret_cast<const int4*>(row + c * 16);
        xx = __dp4a(x.x, x.x, xx); xx = __dp4a(x.y, x.y, xx); xx = __dp4a(x.z, x.z, xx); xx = __dp4a(x.w, x.w, xx);
#pragma unroll
        for (int i = 0; i < NQ; i++) {
            int4 y = *reinterpret_cast<const int4*>(q + i * q_stride + c * 16);
            int32_t d = dot[i];
            d = __dp4a(x.x, y.x, d); d = __dp4a(x.y, y.y, d); d = __dp4a(x.z, y.z, d); d = __dp4a(x.w, y.w, d);
            dot[i] = d;
        }
    }
}

// Per-query scalars prepared once per query.
struct QueryScalars {
    float qnorm;     // f32 cosine: sqrt(sequential sum q^2)  (src/cpu/exhaustive.rs:168-172);
                     // bf16 self-query: that norm rounded to bf16 and widened (exhaustive_bf16.rs:259-270)
    int32_t qnorm_sq;  // sq8: sum of squared codes
};

// Final distance from the raw accumulation.
//   f32/bf16 cosine: 1 - dot / (qnorm * xnorm)        src/utils/dist.rs:3070-3075, 4846-4857
//   prenorm cosine : 1 - dot                          src/utils/k_means_utils.rs:111-133
template <int MET>
__device__ __forceinline__ float finish_fp(float raw, float qnorm, float xnorm) {
    if (MET == MET_L2 || MET == MET_DOT) return raw;
    if (MET == MET_COS_PRENORM) return __fsub_rn(1.0f, raw);
    return __fsub_rn(1.0f, __fdiv_rn(raw, __fmul_rn(qnorm, xnorm)));
}
//   sq8 L2    : sum (q - x)^2 = q.q - 2 q.x + x.x as i32 -> f32     src/utils/dist.rs:5015-5037
//   sq8 cosine: 1 - dot / (sqrt(qn) * sqrt(xn)), 1 if a norm is 0   src/utils/dist.rs:5040-5076
template <int MET>
__device__ __forceinline__ float finish_i8(int32_t dot, int32_t xx, int32_t qn, int32_t xn_stored) {
    if (MET == MET_L2) return __int2float_rn(qn - 2 * dot + xx);
    float a = __fsqrt_rn(__int2float_rn(qn));
    float b = __fsqrt_rn(__int2float_rn(xn_stored));
    if (a > 0.0f && b > 0.0f) return __fsub_rn(1.0f, __fdiv_rn(__int2float_rn(dot), __fmul_rn(a, b)));
    return 1.0f;
}

// Sequential-fold query norm (one thread): src/cpu/exhaustive.rs:168-172, src/cpu/ivf.rs:349-357.
template <int QELEM>
__device__ __forceinline__ float seq_norm(const uint8_t* q, int dim) {
    float s = 0.0f;
    for (int e = 0; e < dim; e++) {
        float v = load1<QELEM>(q, e);
        s = __fadd_rn(s, __fmul_rn(v, v));
    }
    return __fsqrt_rn(s);
}
// f32 -> bf16 -> f32 (RNE, as half::bf16::from_f32).
__device__ __forceinline__ float round_to_bf16(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }

// calculate_l2_norm / dot_simd(v, v) in the AVX2 lane order (rows in global memory).
__device__ __forceinline__ float ref_dot_self_f32(const float* v, int dim) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.0f;
    int chunks = dim >> 3;
    for (int c = 0; c < chunks; c++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float x = v[c * 8 + j];
            acc[j] = __fadd_rn(acc[j], __fmul_rn(x, x));
        }
    float sum = hsum_wide(acc);
    for (int e = chunks * 8; e < dim; e++) sum = __fadd_rn(sum, __fmul_rn(v[e], v[e]));
    return sum;
}

}  // namespace annb
