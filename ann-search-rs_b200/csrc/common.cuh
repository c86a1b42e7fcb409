// common.cuh -- shared device/host helpers for libannb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/annb200.h"

namespace annb {

// ---------------------------------------------------------------------------
// error plumbing: every CUDA failure becomes ANNB_ERR_CUDA + thread-local text
// ---------------------------------------------------------------------------
void set_last_error(const std::string& msg);

#define ANNB_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            ::annb::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) +    \
                                   " (" __FILE__ ":" + std::to_string(__LINE__) + ")");    \
            return ANNB_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

#define ANNB_TRY(expr)                  \
    do {                                \
        int _s = (expr);                \
        if (_s != ANNB_OK) return _s;   \
    } while (0)

// ---------------------------------------------------------------------------
// Candidate keys.  A candidate is (distance, index) ordered by distance, then
// index -- the total order of the reference's (OrderedFloat, usize) tuples
// (src/utils/heap_structs.rs:12-38, 115-132) and of its GPU top-k spec
// (src/gpu/topk_gpu.rs:9-15).  Packed as one u64 so a single unsigned compare
// orders them: high word = monotone map of the f32 bits, low word = index.
// ---------------------------------------------------------------------------
constexpr uint64_t KEY_SENTINEL = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t IDX_INVALID = 0xFFFFFFFFu;

__host__ __device__ __forceinline__ uint32_t f32_to_ordered(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f + 0.0f);  // -0.0 -> +0.0 (reference treats them as equal)
#else
    float g = f + 0.0f;
    uint32_t b;
    memcpy(&b, &g, 4);
#endif
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_f32(uint32_t u) {
    uint32_t b = u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu);
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float dist, uint32_t idx) {
    return (static_cast<uint64_t>(f32_to_ordered(dist)) << 32) | idx;
}
__host__ __device__ __forceinline__ uint32_t key_idx(uint64_t k) { return static_cast<uint32_t>(k); }
__host__ __device__ __forceinline__ float key_dist(uint64_t k) { return ordered_to_f32(static_cast<uint32_t>(k >> 32)); }

__host__ __device__ __forceinline__ uint32_t next_pow2(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}
template <typename T>
__host__ __device__ __forceinline__ T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <typename T>
__host__ __device__ __forceinline__ T round_up(T a, T b) { return ceil_div(a, b) * b; }

// Row storage: every stored row is padded with zeros to a multiple of 16 bytes
// so that rows can be moved with 128-bit loads / cp.async / TMA.
__host__ __device__ __forceinline__ uint32_t elem_bytes(int dtype) { return dtype == ANNB_F32 ? 4u : (dtype == ANNB_BF16 ? 2u : 1u); }
__host__ __device__ __forceinline__ uint32_t padded_row_bytes(uint32_t dim, int dtype) { return round_up(dim * elem_bytes(dtype), 16u); }

// cp.async helpers (LDGSTS), 16-byte granularity.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

}  // namespace annb
