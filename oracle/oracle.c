/*
 * oracle.c -- CPU restatement of ann-search-rs's flat / IVF kNN hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or
 * executed by the product path (ann-search-rs_b200/, include/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker / the timed CPU baseline.
 *
 * Parity status: the reference is a Rust crate and no Rust toolchain exists in
 * this image, so the reference itself cannot be executed here.  This file is
 * pinned against the golden vectors / known-answer tests held by the
 * reference's own unit tests (tests/test_oracle_kat.py lists each with its
 * file:line).  Three third-party details are restated from their published
 * algorithms and are NOT pinned bit-for-bit ("parity unpinned" for them):
 *   - wide 1.4.0  f32x8::reduce_add horizontal-add order (AVX build assumed:
 *     (a0+a4, a1+a5, a2+a6, a3+a7) -> (s0+s2, s1+s3) -> t0+t1);
 *   - faer 0.23.2 matmul rounding inside gemm_assign (dim >= 96): this oracle
 *     uses the direct_assign arithmetic for every dim;
 *   - rand 0.9.4 StdRng streams (k-means sampling / init, synthetic data).
 * Everything else follows the reference source line by line; each function
 * cites the lines it restates (paths relative to /root/reference).
 *
 * Build: see oracle/Makefile (gcc -O3 -mavx2 -mfma -ffp-contract=off -fopenmp).
 * -ffp-contract=off matters: the f32 kernels of the reference use separate
 * multiply and add (wide's `acc += d * d`), the bf16 kernels use explicit FMA.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <immintrin.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_L2 0
#define ORC_COSINE 1
#define ORC_F32 0
#define ORC_BF16 1
#define ORC_SQ8 2

/* ------------------------------------------------------------------------ */
/* f32 SIMD kernels: src/utils/dist.rs:306-330 (euclidean_f32_avx2),         */
/* :587-609 (dot_f32_avx2), :2339-2360 (compute_l2_norm_f32_avx2).           */
/* The runtime dispatch (dist.rs:2786-2805) picks the AVX2 body on any x86   */
/* with AVX2, and also for SimdLevel::Avx512 unless the crate was compiled   */
/* with target_feature=avx512f (dist.rs:371-376).                            */
/* ------------------------------------------------------------------------ */

/* wide 1.4.0 f32x8::reduce_add (AVX path). */
static inline float hsum_wide(__m256 v) {
    __m128 hi = _mm256_extractf128_ps(v, 1);
    __m128 lo = _mm256_castps256_ps128(v);
    __m128 s = _mm_add_ps(lo, hi);            /* a0+a4 a1+a5 a2+a6 a3+a7 */
    __m128 h = _mm_movehl_ps(s, s);           /* s2 s3 . .               */
    __m128 t = _mm_add_ps(s, h);              /* s0+s2 s1+s3             */
    __m128 u = _mm_shuffle_ps(t, t, 0x1);     /* t1                      */
    return _mm_cvtss_f32(_mm_add_ss(t, u));   /* t0 + t1                 */
}

float orc_euclid_f32(const float* a, const float* b, int len) {
    int chunks = len / 8;
    __m256 acc = _mm256_setzero_ps();
    for (int i = 0; i < chunks; i++) {
        __m256 va = _mm256_loadu_ps(a + i * 8);
        __m256 vb = _mm256_loadu_ps(b + i * 8);
        __m256 d = _mm256_sub_ps(va, vb);
        acc = _mm256_add_ps(acc, _mm256_mul_ps(d, d));
    }
    float sum = hsum_wide(acc);
    for (int i = chunks * 8; i < len; i++) {
        float d = a[i] - b[i];
        sum += d * d;
    }
    return sum;
}

float orc_dot_f32(const float* a, const float* b, int len) {
    int chunks = len / 8;
    __m256 acc = _mm256_setzero_ps();
    for (int i = 0; i < chunks; i++) {
        __m256 va = _mm256_loadu_ps(a + i * 8);
        __m256 vb = _mm256_loadu_ps(b + i * 8);
        acc = _mm256_add_ps(acc, _mm256_mul_ps(va, vb));
    }
    float sum = hsum_wide(acc);
    for (int i = chunks * 8; i < len; i++) sum += a[i] * b[i];
    return sum;
}

float orc_l2_norm_f32(const float* v, int len) {
    int chunks = len / 8;
    __m256 acc = _mm256_setzero_ps();
    for (int i = 0; i < chunks; i++) {
        __m256 x = _mm256_loadu_ps(v + i * 8);
        acc = _mm256_add_ps(acc, _mm256_mul_ps(x, x));
    }
    float sum = hsum_wide(acc);
    for (int i = chunks * 8; i < len; i++) sum += v[i] * v[i];
    return sqrtf(sum);
}

/* Plain sequential fold used for query norms:
 * src/cpu/exhaustive.rs:168-172, src/cpu/ivf.rs:349-357. */
float orc_seq_norm_f32(const float* v, int len) {
    float s = 0.0f;
    for (int i = 0; i < len; i++) s = s + v[i] * v[i];
    return sqrtf(s);
}

/* src/utils/dist.rs:5336-5344 normalise_vector. */
void orc_normalise_f32(float* v, int len) {
    float n = orc_l2_norm_f32(v, len);
    if (n > 0.0f)
        for (int i = 0; i < len; i++) v[i] = v[i] / n;
}

/* Row-wise helpers (index constructors: exhaustive.rs:86-96, ivf_sq8.rs:169-176). */
void orc_row_norms_f32(const float* x, int64_t n, int dim, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) out[i] = orc_l2_norm_f32(x + i * dim, dim);
}
void orc_normalise_rows_f32(float* x, int64_t n, int dim) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) orc_normalise_f32(x + i * dim, dim);
}

/* ------------------------------------------------------------------------ */
/* BF16: src/quantised/quantisers.rs:31-38 (encode, half 2.7.1 RNE),         */
/* src/utils/dist.rs:3198-3209 (widen = <<16), :3167-3178 / :3211-3220      */
/* (hsum), :3392-3416, :3615-3636, :4118-4150, :4322-4357 (AVX2 kernels).   */
/* ------------------------------------------------------------------------ */

uint16_t orc_f32_to_bf16(float value) {
    uint32_t x;
    memcpy(&x, &value, 4);
    if ((x & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((x >> 16) | 0x0040u);
    uint32_t round_bit = 0x00008000u;
    if ((x & round_bit) != 0 && (x & (3 * round_bit - 1)) != 0) return (uint16_t)((x >> 16) + 1);
    return (uint16_t)(x >> 16);
}

static inline float bf16_to_f32(uint16_t h) {
    uint32_t x = ((uint32_t)h) << 16;
    float f;
    memcpy(&f, &x, 4);
    return f;
}
float orc_bf16_to_f32(uint16_t h) { return bf16_to_f32(h); }

void orc_encode_bf16(const float* src, uint16_t* dst, int64_t len) {
    for (int64_t i = 0; i < len; i++) dst[i] = orc_f32_to_bf16(src[i]);
}

static inline __m256 bf16x8(const uint16_t* p) {
    __m128i raw = _mm_loadu_si128((const __m128i*)p);
    return _mm256_castsi256_ps(_mm256_slli_epi32(_mm256_cvtepu16_epi32(raw), 16));
}

/* hsum_f32_avx2 -> hsum_f32_sse (dist.rs:3167-3178, 3211-3220). */
static inline float hsum_bf16path(__m256 v) {
    __m128 low = _mm256_castps256_ps128(v);
    __m128 high = _mm256_extractf128_ps(v, 1);
    __m128 s = _mm_add_ps(low, high);
    __m128 shuf = _mm_movehdup_ps(s);
    __m128 sums = _mm_add_ps(s, shuf);        /* s0+s1 . s2+s3 . */
    __m128 shuf2 = _mm_movehl_ps(sums, sums);
    return _mm_cvtss_f32(_mm_add_ss(sums, shuf2));
}

float orc_euclid_bf16_f32(const uint16_t* a, const float* b, int len) {
    int chunks = len / 8;
    __m256 acc = _mm256_setzero_ps();
    for (int i = 0; i < chunks; i++) {
        __m256 d = _mm256_sub_ps(bf16x8(a + i * 8), _mm256_loadu_ps(b + i * 8));
        acc = _mm256_fmadd_ps(d, d, acc);
    }
    float sum = hsum_bf16path(acc);
    for (int i = chunks * 8; i < len; i++) {
        float d = bf16_to_f32(a[i]) - b[i];
        sum += d * d;
    }
    return sum;
}

float orc_dot_bf16_f32(const uint16_t* a, const float* b, int len) {
    int chunks = len / 8;
    __m256 acc = _mm256_setzero_ps();
    for (int i = 0; i < chunks; i++)
        acc = _mm256_fmadd_ps(bf16x8(a + i * 8), _mm256_loadu_ps(b + i * 8), acc);
    float sum = hsum_bf16path(acc);
    for (int i = chunks * 8; i < len; i++) sum += bf16_to_f32(a[i]) * b[i];
    return sum;
}

float orc_euclid_bf16_bf16(const uint16_t* a, const uint16_t* b, int len) {
    int chunks = len / 8;
    __m256 acc = _mm256_setzero_ps();
    for (int i = 0; i < chunks; i++) {
        __m256 d = _mm256_sub_ps(bf16x8(a + i * 8), bf16x8(b + i * 8));
        acc = _mm256_fmadd_ps(d, d, acc);
    }
    float sum = hsum_bf16path(acc);
    for (int i = chunks * 8; i < len; i++) {
        float d = bf16_to_f32(a[i]) - bf16_to_f32(b[i]);
        sum += d * d;
    }
    return sum;
}

float orc_dot_bf16_bf16(const uint16_t* a, const uint16_t* b, int len) {
    int chunks = len / 8;
    __m256 acc = _mm256_setzero_ps();
    for (int i = 0; i < chunks; i++)
        acc = _mm256_fmadd_ps(bf16x8(a + i * 8), bf16x8(b + i * 8), acc);
    float sum = hsum_bf16path(acc);
    for (int i = chunks * 8; i < len; i++) sum += bf16_to_f32(a[i]) * bf16_to_f32(b[i]);
    return sum;
}

/* src/quantised/quantisers.rs:80-91 bf16_norm (sequential fold). */
float orc_bf16_norm(const uint16_t* v, int len) {
    float s = 0.0f;
    for (int i = 0; i < len; i++) {
        float f = bf16_to_f32(v[i]);
        s = s + f * f;
    }
    return sqrtf(s);
}

/* ------------------------------------------------------------------------ */
/* SQ8: src/quantised/quantisers.rs:123-183, src/utils/dist.rs:5015-5077.    */
/* ------------------------------------------------------------------------ */

void orc_sq8_train(const float* data, int64_t n, int dim, float* scales) {
    for (int d = 0; d < dim; d++) {
        float mx = 0.0f;
        for (int64_t i = 0; i < n; i++) {
            float a = fabsf(data[i * dim + d]);
            mx = (a > mx) ? a : mx; /* Float::max ignores NaN operands */
        }
        scales[d] = (mx <= 0.0f) ? 1.0f : mx / 128.0f;
    }
}

static inline int8_t sq8_encode_one(float val, float scale) {
    float scaled = val / scale;
    float sg = signbit(scaled) ? -1.0f : 1.0f; /* f32::signum: +0 -> 1, -0 -> -1 */
    if (isnan(scaled)) return 0;
    float rounded = scaled + 0.5f * sg;
    float clamped = fminf(rounded, 127.0f);
    clamped = fmaxf(clamped, -128.0f);
    return (int8_t)clamped; /* trunc toward zero, always in range */
}

void orc_sq8_encode(const float* vec, const float* scales, int dim, int8_t* out) {
    for (int d = 0; d < dim; d++) out[d] = sq8_encode_one(vec[d], scales[d]);
}

void orc_sq8_encode_all(const float* data, int64_t n, int dim, const float* scales, int8_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) orc_sq8_encode(data + i * dim, scales, dim, out + i * dim);
}

void orc_sq8_decode(const int8_t* q, const float* scales, int dim, float* out) {
    for (int d = 0; d < dim; d++) out[d] = (float)q[d] * scales[d];
}

int32_t orc_sq8_norm_sq(const int8_t* v, int dim) {
    int32_t s = 0;
    for (int d = 0; d < dim; d++) s += (int32_t)v[d] * (int32_t)v[d];
    return s;
}

float orc_sq8_euclid(const int8_t* db, const int8_t* q, int dim) {
    int32_t sum = 0;
    for (int d = 0; d < dim; d++) {
        int32_t diff = (int32_t)q[d] - (int32_t)db[d];
        sum += diff * diff;
    }
    return (float)sum;
}

float orc_sq8_cosine(const int8_t* db, int32_t db_norm_sq, const int8_t* q, int32_t q_norm_sq, int dim) {
    int32_t dot = 0;
    for (int d = 0; d < dim; d++) dot += (int32_t)q[d] * (int32_t)db[d];
    float qn = sqrtf((float)q_norm_sq);
    float dn = sqrtf((float)db_norm_sq);
    if (qn > 0.0f && dn > 0.0f) return 1.0f - (float)dot / (qn * dn);
    return 1.0f;
}

/* ------------------------------------------------------------------------ */
/* Result containers.                                                        */
/* ------------------------------------------------------------------------ */

typedef struct {
    float d;
    int64_t i;
} pair_t;

/* Total order of (OrderedFloat<T>, usize): src/utils/heap_structs.rs:12-38
 * (NaN compares Equal), then the index. */
static inline int pair_less(pair_t a, pair_t b) {
    if (a.d < b.d) return 1;
    if (a.d > b.d) return 0;
    return a.i < b.i;
}

/* std BinaryHeap<(OrderedFloat, usize)> restated as an explicit max-heap. */
typedef struct {
    pair_t* a;
    int len;
} heap_t;

static void heap_push(heap_t* h, pair_t v) {
    int i = h->len++;
    h->a[i] = v;
    while (i > 0) {
        int p = (i - 1) / 2;
        if (pair_less(h->a[p], h->a[i])) {
            pair_t t = h->a[p];
            h->a[p] = h->a[i];
            h->a[i] = t;
            i = p;
        } else
            break;
    }
}

static void heap_pop(heap_t* h) {
    h->a[0] = h->a[--h->len];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < h->len && pair_less(h->a[m], h->a[l])) m = l;
        if (r < h->len && pair_less(h->a[m], h->a[r])) m = r;
        if (m == i) break;
        pair_t t = h->a[m];
        h->a[m] = h->a[i];
        h->a[i] = t;
        i = m;
    }
}

/* The scan-time rule shared by every heap user, e.g. src/cpu/exhaustive.rs:156-165:
 * push while len < k, else replace the max only on a strictly smaller distance. */
static inline void heap_offer(heap_t* h, int k, float dist, int64_t idx) {
    if (k <= 0) return;
    if (h->len < k) {
        pair_t p = {dist, idx};
        heap_push(h, p);
    } else if (dist < h->a[0].d) {
        heap_pop(h);
        pair_t p = {dist, idx};
        heap_push(h, p);
    }
}

static int pair_cmp_qsort(const void* x, const void* y) {
    pair_t a = *(const pair_t*)x, b = *(const pair_t*)y;
    if (pair_less(a, b)) return -1;
    if (pair_less(b, a)) return 1;
    return 0;
}

/* The reference finishes with sort_unstable_by_key(dist) (exhaustive.rs:199-200):
 * order inside an equal-distance run is unspecified there; the oracle fixes it
 * to ascending index so that outputs are deterministic. */
static void heap_finish(heap_t* h) { qsort(h->a, h->len, sizeof(pair_t), pair_cmp_qsort); }

/* SortedBuffer::insert, src/utils/heap_structs.rs:115-132. */
static inline void sorted_insert(pair_t* buf, int* len, int limit, pair_t item) {
    if (limit <= 0) return;
    if (*len < limit) {
        int lo = 0, hi = *len;
        while (lo < hi) {
            int mid = (lo + hi) / 2;
            if (pair_less(buf[mid], item)) lo = mid + 1; else hi = mid;
        }
        memmove(buf + lo + 1, buf + lo, (size_t)(*len - lo) * sizeof(pair_t));
        buf[lo] = item;
        (*len)++;
    } else if (pair_less(item, buf[*len - 1])) {
        int lo = 0, hi = *len;
        while (lo < hi) {
            int mid = (lo + hi) / 2;
            if (pair_less(buf[mid], item)) lo = mid + 1; else hi = mid;
        }
        memmove(buf + lo + 1, buf + lo, (size_t)(*len - 1 - lo) * sizeof(pair_t));
        buf[lo] = item;
    }
}

/* ------------------------------------------------------------------------ */
/* Flat index view + per-pair distance, all dtypes.                          */
/* ------------------------------------------------------------------------ */

typedef struct {
    int dtype, metric, dim;
    int64_t n;
    const void* vectors;     /* f32 / bf16(u16) / i8 row-major              */
    const float* norms;      /* f32, bf16 cosine: f32 norms of original rows */
    const int32_t* norms_i;  /* sq8 cosine: sum code^2                       */
} store_t;

/* External f32 query against one stored row.
 * f32 : dist.rs:3026-3030, 3070-3075
 * bf16: dist.rs:4805-4813, 4846-4857
 * sq8 : handled by the caller (query is encoded first). */
static inline float dist_f32q(const store_t* s, int64_t idx, const float* q, float qnorm) {
    if (s->dtype == ORC_F32) {
        const float* v = (const float*)s->vectors + idx * s->dim;
        if (s->metric == ORC_L2) return orc_euclid_f32(v, q, s->dim);
        float dot = orc_dot_f32(v, q, s->dim);
        return 1.0f - (dot / (qnorm * s->norms[idx]));
    } else {
        const uint16_t* v = (const uint16_t*)s->vectors + idx * s->dim;
        if (s->metric == ORC_L2) return orc_euclid_bf16_f32(v, q, s->dim);
        float dot = orc_dot_bf16_f32(v, q, s->dim);
        return 1.0f - (dot / (qnorm * s->norms[idx]));
    }
}

/* Self-query of a bf16 row (dual bf16): dist.rs:4815-4823, 4859-4875;
 * the query norm is rounded to bf16 (exhaustive_bf16.rs:259-270). */
static inline float dist_bf16q(const store_t* s, int64_t idx, const uint16_t* q, uint16_t qnorm_bf16) {
    const uint16_t* v = (const uint16_t*)s->vectors + idx * s->dim;
    if (s->metric == ORC_L2) return orc_euclid_bf16_bf16(v, q, s->dim);
    float dot = orc_dot_bf16_bf16(v, q, s->dim);
    return 1.0f - (dot / (bf16_to_f32(qnorm_bf16) * s->norms[idx]));
}

static inline float dist_i8q(const store_t* s, int64_t idx, const int8_t* q, int32_t qn) {
    const int8_t* v = (const int8_t*)s->vectors + idx * s->dim;
    if (s->metric == ORC_L2) return orc_sq8_euclid(v, q, s->dim);
    return orc_sq8_cosine(v, s->norms_i[idx], q, qn, s->dim);
}

/* ------------------------------------------------------------------------ */
/* Flat query.  f32: src/cpu/exhaustive.rs:142-208; bf16:                     */
/* src/quantised/exhaustive_bf16.rs:142-197 (query), 239-292 (query_bf16);   */
/* sq8: src/quantised/exhaustive_sq8.rs:172-230 (query), 277-340 (self).     */
/* `self_mode` = generate_knn: query i is stored row `self_rows[i]`.         */
/* Outputs are padded to k with id = -1, dist = +inf; counts[i] = min(k, n). */
/* ------------------------------------------------------------------------ */
int orc_flat_search(int dtype, int metric, const void* vectors, int64_t n, int dim,
                    const float* norms, const int32_t* norms_i, const float* sq8_scales,
                    const float* queries, int64_t nq, const int64_t* self_rows, int self_mode,
                    int k, int64_t* out_ids, float* out_dist, int32_t* out_counts, int nthreads) {
    store_t s = {dtype, metric, dim, n, vectors, norms, norms_i};
    int kk = (int)((int64_t)k < n ? k : n);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        pair_t* hbuf = (pair_t*)malloc(sizeof(pair_t) * (size_t)(kk + 1));
        float* qtmp = (float*)malloc(sizeof(float) * (size_t)dim);
        int8_t* qi8 = (int8_t*)malloc((size_t)dim);
#pragma omp for schedule(dynamic, 4)
        for (int64_t qi = 0; qi < nq; qi++) {
            heap_t h = {hbuf, 0};
            if (self_mode) {
                int64_t row = self_rows ? self_rows[qi] : qi;
                if (dtype == ORC_F32) {
                    const float* q = (const float*)vectors + row * dim;
                    float qn = (metric == ORC_COSINE) ? orc_seq_norm_f32(q, dim) : 1.0f;
                    for (int64_t i = 0; i < n; i++) heap_offer(&h, kk, dist_f32q(&s, i, q, qn), i);
                } else if (dtype == ORC_BF16) {
                    const uint16_t* q = (const uint16_t*)vectors + row * dim;
                    uint16_t qn = (metric == ORC_COSINE) ? orc_f32_to_bf16(orc_bf16_norm(q, dim)) : 0;
                    for (int64_t i = 0; i < n; i++) heap_offer(&h, kk, dist_bf16q(&s, i, q, qn), i);
                } else {
                    const int8_t* q = (const int8_t*)vectors + row * dim;
                    int32_t qn = (metric == ORC_COSINE) ? norms_i[row] : 0;
                    for (int64_t i = 0; i < n; i++) heap_offer(&h, kk, dist_i8q(&s, i, q, qn), i);
                }
            } else {
                const float* q = queries + qi * dim;
                if (dtype == ORC_SQ8) {
                    memcpy(qtmp, q, sizeof(float) * (size_t)dim);
                    if (metric == ORC_COSINE) orc_normalise_f32(qtmp, dim);
                    orc_sq8_encode(qtmp, sq8_scales, dim, qi8);
                    int32_t qn = orc_sq8_norm_sq(qi8, dim);
                    for (int64_t i = 0; i < n; i++) heap_offer(&h, kk, dist_i8q(&s, i, qi8, qn), i);
                } else {
                    float qn = (metric == ORC_COSINE) ? orc_seq_norm_f32(q, dim) : 1.0f;
                    for (int64_t i = 0; i < n; i++) heap_offer(&h, kk, dist_f32q(&s, i, q, qn), i);
                }
            }
            heap_finish(&h);
            for (int j = 0; j < k; j++) {
                out_ids[qi * k + j] = (j < h.len) ? h.a[j].i : -1;
                if (out_dist) out_dist[qi * k + j] = (j < h.len) ? h.a[j].d : INFINITY;
            }
            if (out_counts) out_counts[qi] = h.len;
        }
        free(hbuf);
        free(qtmp);
        free(qi8);
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* IVF pieces.                                                               */
/* ------------------------------------------------------------------------ */

/* build_csr_layout, src/utils/k_means_utils.rs:2955-2980. */
void orc_build_csr(const int64_t* assign, int64_t n, int nlist, int64_t* all_indices, int64_t* offsets) {
    for (int i = 0; i <= nlist; i++) offsets[i] = 0;
    for (int64_t i = 0; i < n; i++) offsets[assign[i] + 1] += 1;
    for (int i = 1; i <= nlist; i++) offsets[i] += offsets[i - 1];
    int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nlist + 1));
    memcpy(cur, offsets, sizeof(int64_t) * (size_t)(nlist + 1));
    for (int64_t i = 0; i < n; i++) all_indices[cur[assign[i]]++] = i;
    free(cur);
}

/* direct_assign, src/utils/k_means_utils.rs:2119-2195 (strict `>`: lowest
 * centroid id wins ties).  Used for every dim (see header note on faer). */
void orc_assign_all(const float* data, int64_t n, int dim, const float* centroids,
                    const float* centroid_norms, int nlist, int metric, int64_t* out, int nthreads) {
    float* sc = (float*)malloc(sizeof(float) * (size_t)nlist);
    for (int c = 0; c < nlist; c++) {
        const float* cent = centroids + (int64_t)c * dim;
        if (metric == ORC_L2) sc[c] = orc_dot_f32(cent, cent, dim);
        else {
            float nm = centroid_norms[c];
            sc[c] = (nm > 0.0f) ? 1.0f / nm : 0.0f;
        }
    }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const float* v = data + i * dim;
        int best = 0;
        float mx = -INFINITY;
        for (int c = 0; c < nlist; c++) {
            const float* cent = centroids + (int64_t)c * dim;
            float score = (metric == ORC_L2) ? 2.0f * orc_dot_f32(v, cent, dim) - sc[c]
                                             : orc_dot_f32(v, cent, dim) * sc[c];
            if (score > mx) {
                mx = score;
                best = c;
            }
        }
        out[i] = best;
    }
    free(sc);
}

typedef struct {
    float d;
    int c;
} cd_t;

static int cd_cmp(const void* x, const void* y) {
    const cd_t* a = (const cd_t*)x;
    const cd_t* b = (const cd_t*)y;
    if (a->d < b->d) return -1;
    if (a->d > b->d) return 1;
    /* sort_unstable_by(dist) leaves equal-distance cells in unspecified order;
     * the oracle fixes ascending cell id. */
    return (a->c > b->c) - (a->c < b->c);
}

/* select_probed_clusters, src/utils/k_means_utils.rs:3007-3029. */
int orc_select_probed(const float* dists, const int32_t* cells, int nlist, const int64_t* offsets,
                      int nprobe, int64_t k, int32_t* chosen) {
    cd_t* cd = (cd_t*)malloc(sizeof(cd_t) * (size_t)nlist);
    for (int i = 0; i < nlist; i++) {
        cd[i].d = dists[i];
        cd[i].c = cells ? cells[i] : i;
    }
    qsort(cd, (size_t)nlist, sizeof(cd_t), cd_cmp);
    int cnt = 0;
    int64_t reach = 0;
    for (int i = 0; i < nlist; i++) {
        int c = cd[i].c;
        chosen[cnt++] = c;
        reach += offsets[c + 1] - offsets[c];
        if (cnt >= nprobe && reach >= k) break;
    }
    free(cd);
    return cnt;
}

/* get_centroids_dist / get_centroids_prenorm,
 * src/utils/k_means_utils.rs:76-99, 111-133.  (The select_nth_unstable step is
 * a no-op for the final result because select_probed_clusters fully sorts.) */
void orc_centroid_dists(const float* q, float qnorm, const float* centroids, const float* cnorms,
                        int nlist, int dim, int metric, int prenorm, float* out) {
    for (int c = 0; c < nlist; c++) {
        const float* cent = centroids + (int64_t)c * dim;
        if (metric == ORC_L2) out[c] = prenorm ? orc_euclid_f32(q, cent, dim) : orc_euclid_f32(q, cent, dim);
        else if (prenorm) out[c] = 1.0f - orc_dot_f32(q, cent, dim);
        else out[c] = 1.0f - (orc_dot_f32(q, cent, dim) / (qnorm * cnorms[c]));
    }
}

/* IVF query, all dtypes.
 * f32 : src/cpu/ivf.rs:337-390 (SortedBuffer keyed (dist, internal idx))
 * bf16: src/quantised/ivf_bf16.rs:277-332 (query), 450-501 (query_bf16)
 * sq8 : src/quantised/ivf_sq8.rs:303-359 (query), 397-444 (query_quantised);
 *       heap instead of SortedBuffer, normalised query, prenorm routing.
 * vectors are in list order; ids returned are original_ids[internal].
 * nprobe <= 0 means "None" -> max(1, floor(sqrt(nlist))).
 * Also reports probed-list count and scanned-vector count per query. */
int orc_ivf_search(int dtype, int metric, const void* vectors, int64_t n, int dim,
                   const float* norms, const int32_t* norms_i, const float* sq8_scales,
                   const float* centroids, const float* centroid_norms, int nlist,
                   const int64_t* offsets, const int64_t* original_ids,
                   const float* queries, int64_t nq, const int64_t* self_rows, int self_mode,
                   int k, int nprobe, int64_t* out_ids, float* out_dist, int32_t* out_counts,
                   int32_t* out_nprobed, int64_t* out_nscanned, int nthreads) {
    store_t s = {dtype, metric, dim, n, vectors, norms, norms_i};
    int np = nprobe > 0 ? nprobe : (int)sqrt((double)nlist);
    if (np < 1) np = 1;
    if (np > nlist) np = nlist;
    int kk = (int)((int64_t)k < n ? k : n);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        pair_t* buf = (pair_t*)malloc(sizeof(pair_t) * (size_t)(kk + 2));
        float* cdist = (float*)malloc(sizeof(float) * (size_t)nlist);
        int32_t* chosen = (int32_t*)malloc(sizeof(int32_t) * (size_t)nlist);
        float* qtmp = (float*)malloc(sizeof(float) * (size_t)dim);
        int8_t* qi8 = (int8_t*)malloc((size_t)dim);
#pragma omp for schedule(dynamic, 4)
        for (int64_t qi = 0; qi < nq; qi++) {
            int64_t row = self_mode ? (self_rows ? self_rows[qi] : qi) : -1;
            const float* qf = NULL;     /* f32 view used for routing */
            const uint16_t* qb = NULL;  /* bf16 self-query           */
            const int8_t* qc = NULL;    /* sq8 code-space query      */
            float qnorm = 1.0f;
            uint16_t qnorm_b = 0;
            int32_t qn_i = 0;
            int prenorm = 0;
            if (dtype == ORC_F32) {
                qf = self_mode ? (const float*)vectors + row * dim : queries + qi * dim;
                if (metric == ORC_COSINE) qnorm = orc_seq_norm_f32(qf, dim);
            } else if (dtype == ORC_BF16) {
                if (self_mode) {
                    qb = (const uint16_t*)vectors + row * dim;
                    for (int d = 0; d < dim; d++) qtmp[d] = bf16_to_f32(qb[d]);
                    qf = qtmp;
                    if (metric == ORC_COSINE) {
                        qnorm = orc_bf16_norm(qb, dim);          /* routing: unrounded */
                        qnorm_b = orc_f32_to_bf16(qnorm);        /* scan: bf16-rounded */
                    }
                } else {
                    qf = queries + qi * dim;
                    if (metric == ORC_COSINE) qnorm = orc_seq_norm_f32(qf, dim);
                }
            } else {
                prenorm = 1;
                if (self_mode) {
                    qc = (const int8_t*)vectors + row * dim;
                    qn_i = (metric == ORC_COSINE) ? norms_i[row] : 0;
                    orc_sq8_decode(qc, sq8_scales, dim, qtmp);
                    qf = qtmp;
                } else {
                    memcpy(qtmp, queries + qi * dim, sizeof(float) * (size_t)dim);
                    if (metric == ORC_COSINE) orc_normalise_f32(qtmp, dim);
                    qf = qtmp;
                    orc_sq8_encode(qtmp, sq8_scales, dim, qi8);
                    qc = qi8;
                    qn_i = orc_sq8_norm_sq(qi8, dim);
                }
            }
            orc_centroid_dists(qf, qnorm, centroids, centroid_norms, nlist, dim, metric, prenorm, cdist);
            int nch = orc_select_probed(cdist, NULL, nlist, offsets, np, kk, chosen);
            int64_t scanned = 0;
            int len = 0;
            heap_t h = {buf, 0};
            for (int p = 0; p < nch; p++) {
                int c = chosen[p];
                for (int64_t v = offsets[c]; v < offsets[c + 1]; v++) {
                    if (dtype == ORC_SQ8) {
                        heap_offer(&h, kk, dist_i8q(&s, v, qc, qn_i), v);
                    } else {
                        float d = (dtype == ORC_BF16 && self_mode) ? dist_bf16q(&s, v, qb, qnorm_b)
                                                                   : dist_f32q(&s, v, qf, qnorm);
                        pair_t it = {d, v};
                        sorted_insert(buf, &len, kk, it);
                    }
                }
                scanned += offsets[c + 1] - offsets[c];
            }
            if (dtype == ORC_SQ8) {
                heap_finish(&h);
                len = h.len;
            }
            /* full generate_knn (self_rows == NULL) scatters by original id (ivf.rs:476-486);
             * a sub-sampled self query keeps the caller's row order. */
            int64_t orow = (self_mode && !self_rows) ? original_ids[row] : qi;
            for (int j = 0; j < k; j++) {
                out_ids[orow * k + j] = (j < len) ? original_ids[buf[j].i] : -1;
                if (out_dist) out_dist[orow * k + j] = (j < len) ? buf[j].d : INFINITY;
            }
            if (out_counts) out_counts[orow] = len;
            if (out_nprobed) out_nprobed[orow] = nch;
            if (out_nscanned) out_nscanned[orow] = scanned;
        }
        free(buf);
        free(cdist);
        free(chosen);
        free(qtmp);
        free(qi8);
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Stand-in centroid trainer (plain Lloyd, deterministic init from evenly    */
/* spaced training rows).  NOT a restatement: the reference's trainer        */
/* (k_means_utils.rs:2771-2938) depends on rand's StdRng stream, which       */
/* cannot be reproduced here.  IVF parity is defined on shared index         */
/* contents (same centroids / offsets / permutation fed to oracle and GPU).  */
/* Empty clusters keep their previous centroid (k_means_utils.rs:1097-1105). */
/* ------------------------------------------------------------------------ */
void orc_kmeans_lloyd(const float* train, int64_t n, int dim, int nlist, int metric, int iters,
                      float* centroids, int nthreads) {
    for (int c = 0; c < nlist; c++) {
        int64_t r = (int64_t)(((__int128)c * n) / nlist);
        memcpy(centroids + (int64_t)c * dim, train + r * dim, sizeof(float) * (size_t)dim);
    }
    int64_t* assign = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    float* cn = (float*)malloc(sizeof(float) * (size_t)nlist);
    double* sums = (double*)malloc(sizeof(double) * (size_t)nlist * dim);
    int64_t* cnt = (int64_t*)malloc(sizeof(int64_t) * (size_t)nlist);
    for (int it = 0; it < iters; it++) {
        for (int c = 0; c < nlist; c++) cn[c] = orc_seq_norm_f32(centroids + (int64_t)c * dim, dim);
        orc_assign_all(train, n, dim, centroids, cn, nlist, metric, assign, nthreads);
        memset(sums, 0, sizeof(double) * (size_t)nlist * dim);
        memset(cnt, 0, sizeof(int64_t) * (size_t)nlist);
        for (int64_t i = 0; i < n; i++) {
            int64_t c = assign[i];
            cnt[c]++;
            for (int d = 0; d < dim; d++) sums[c * dim + d] += train[i * dim + d];
        }
        for (int c = 0; c < nlist; c++)
            if (cnt[c] > 0)
                for (int d = 0; d < dim; d++) centroids[(int64_t)c * dim + d] = (float)(sums[(int64_t)c * dim + d] / (double)cnt[c]);
    }
    free(assign);
    free(cn);
    free(sums);
    free(cnt);
}

/* parallel_lloyd, unbalanced (src/utils/k_means_utils.rs:1572-1700), from caller-supplied initial centroids:
 * assign (direct_assign; cosine with calculate_l2_norm of the current centroids) -> stop when at most
 * max(1, n / 10000) assignments changed, tested BEFORE the update (:1611-1624) -> centroid = sum / count, empty
 * clusters keep their centroid (:1657-1666).  The reference sums f32 partials per rayon chunk (order depends on the
 * pool size, not reproducible); this restatement and the device kernel accumulate in f64, the comparison between
 * them is by tolerance.  Returns the number of centroid updates performed. */
int orc_parallel_lloyd(const float* data, int64_t n, int dim, int nlist, int metric, int max_iters, float* centroids,
                       int nthreads) {
    int64_t* assign = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    int64_t* prev = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    float* cn = (float*)malloc(sizeof(float) * (size_t)nlist);
    double* sums = (double*)malloc(sizeof(double) * (size_t)nlist * dim);
    int64_t* cnt = (int64_t*)malloc(sizeof(int64_t) * (size_t)nlist);
    for (int64_t i = 0; i < n; i++) prev[i] = -1;
    const int64_t change_floor = (n / 10000) > 1 ? (n / 10000) : 1;
    int it = 0;
    for (; it < max_iters; it++) {
        for (int c = 0; c < nlist; c++) cn[c] = orc_l2_norm_f32(centroids + (int64_t)c * dim, dim);
        orc_assign_all(data, n, dim, centroids, cn, nlist, metric, assign, nthreads);
        int64_t changed = 0;
        for (int64_t i = 0; i < n; i++) changed += assign[i] != prev[i];
        if (changed <= change_floor) break;
        memset(sums, 0, sizeof(double) * (size_t)nlist * dim);
        memset(cnt, 0, sizeof(int64_t) * (size_t)nlist);
        for (int64_t i = 0; i < n; i++) {
            int64_t c = assign[i];
            cnt[c]++;
            for (int d = 0; d < dim; d++) sums[c * dim + d] += data[i * dim + d];
        }
        for (int c = 0; c < nlist; c++)
            if (cnt[c] > 0)
                for (int d = 0; d < dim; d++) centroids[(int64_t)c * dim + d] = (float)(sums[(int64_t)c * dim + d] / (double)cnt[c]);
        memcpy(prev, assign, sizeof(int64_t) * (size_t)n);
    }
    free(assign);
    free(prev);
    free(cn);
    free(sums);
    free(cnt);
    return it;
}

/* adjust_centers (src/utils/k_means_utils.rs:979-1030), the balancing step of RAFT's balanced k-means: every centroid
 * whose cluster holds at most BALANCE_THRESHOLD (0.25) of the average size is pulled toward a point of an above-average
 * cluster, weight min(count, BALANCE_PULLBACK = 5); the donor is found by a strided walk (BALANCE_DONOR_STRIDE =
 * 715827883) whose cursor carries over from one starved centroid to the next.  No random numbers: fully reproducible.
 * Returns the number of centroids moved. */
int64_t orc_adjust_centers(float* centroids, int dim, int k, const float* data, int64_t n, const int64_t* assign, const int64_t* counts,
                           uint64_t seed) {
    if (k == 0 || n == 0) return 0;
    const double average = (double)n / (double)k;
    const double floor_ = average * 0.25;
    int64_t adjusted = 0;
    uint64_t cursor = seed % (uint64_t)n;
    for (int c = 0; c < k; c++) {
        if ((double)counts[c] > floor_) continue;
        int64_t donor = -1;
        for (int64_t t = 0; t < n; t++) {
            cursor = (cursor + 715827883ull) % (uint64_t)n;
            const int64_t owner = assign[cursor];
            if (owner != c && (double)counts[owner] > average) { donor = (int64_t)cursor; break; }
        }
        if (donor < 0) continue;
        const float w = (float)(counts[c] < 5 ? counts[c] : 5);
        const float denom = w + 1.0f;
        for (int d = 0; d < dim; d++) {
            float* cc = centroids + (int64_t)c * dim + d;
            *cc = (*cc * w + data[donor * dim + d]) / denom;
        }
        adjusted++;
    }
    return adjusted;
}

/* parallel_lloyd with the balancing hook (src/utils/k_means_utils.rs:1572-1700): as orc_parallel_lloyd, plus adjust_centers
 * with seed + iteration right after the means (:1668-1682), and the loop stays in until balancing has nothing left to do
 * (`changed <= change_floor && last_adjusted == 0`, :1618).  *out_adjusted: total number of centroid moves. */
int orc_parallel_lloyd_balanced(const float* data, int64_t n, int dim, int nlist, int metric, int max_iters, int balanced, uint64_t seed,
                                float* centroids, int64_t* out_adjusted, int nthreads) {
    int64_t* assign = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    int64_t* prev = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    float* cn = (float*)malloc(sizeof(float) * (size_t)nlist);
    double* sums = (double*)malloc(sizeof(double) * (size_t)nlist * dim);
    int64_t* cnt = (int64_t*)malloc(sizeof(int64_t) * (size_t)nlist);
    for (int64_t i = 0; i < n; i++) prev[i] = -1;
    const int64_t change_floor = (n / 10000) > 1 ? (n / 10000) : 1;
    int64_t last_adjusted = 0, total = 0;
    int it = 0;
    for (; it < max_iters; it++) {
        for (int c = 0; c < nlist; c++) cn[c] = orc_l2_norm_f32(centroids + (int64_t)c * dim, dim);
        orc_assign_all(data, n, dim, centroids, cn, nlist, metric, assign, nthreads);
        int64_t changed = 0;
        for (int64_t i = 0; i < n; i++) changed += assign[i] != prev[i];
        if (changed <= change_floor && last_adjusted == 0) break;
        memset(sums, 0, sizeof(double) * (size_t)nlist * dim);
        memset(cnt, 0, sizeof(int64_t) * (size_t)nlist);
        for (int64_t i = 0; i < n; i++) {
            int64_t c = assign[i];
            cnt[c]++;
            for (int d = 0; d < dim; d++) sums[c * dim + d] += data[i * dim + d];
        }
        for (int c = 0; c < nlist; c++)
            if (cnt[c] > 0)
                for (int d = 0; d < dim; d++) centroids[(int64_t)c * dim + d] = (float)(sums[(int64_t)c * dim + d] / (double)cnt[c]);
        if (balanced) {
            last_adjusted = orc_adjust_centers(centroids, dim, nlist, data, n, assign, cnt, seed + (uint64_t)it);
            total += last_adjusted;
        }
        memcpy(prev, assign, sizeof(int64_t) * (size_t)n);
    }
    if (out_adjusted) *out_adjusted = total;
    free(assign);
    free(prev);
    free(cn);
    free(sums);
    free(cnt);
    return it;
}

/* Host-thread count the batch entry points will use. */
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
