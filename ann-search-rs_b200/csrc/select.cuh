// select.cuh -- top-k selection primitives on packed (distance, index) keys.
//
// WarpSelect: one warp keeps the k best keys of a stream.  Keys that beat the
// current k-th key are appended to a shared-memory queue with a ballot-ranked
// write; when the queue fills, the warp bitonic-sorts [best k | queue] in place
// and tightens the threshold.  After the warm-up almost every candidate fails
// the register compare, so the steady-state cost is one compare + one ballot.
// The order is the reference's total order (distance, index), so the result is
// independent of the order in which candidates are offered.
#pragma once
#include "common.cuh"

namespace annb {

// In-place ascending bitonic sort of n (power of two) keys by `nthreads`
// cooperating threads; SYNC is __syncwarp or __syncthreads.
template <bool BLOCK>
__device__ __forceinline__ void bitonic_sort_keys(uint64_t* buf, uint32_t n, uint32_t tid, uint32_t nthreads) {
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = tid; t < (n >> 1); t += nthreads) {
                const uint32_t lowmask = stride - 1u;                          // stride is a power of two: no division
                uint32_t i = ((t & ~lowmask) << 1) | (t & lowmask);
                uint32_t j = i + stride;
                bool up = ((i & size) == 0);
                uint64_t a = buf[i], b = buf[j];
                if ((a > b) == up) {
                    buf[i] = b;
                    buf[j] = a;
                }
            }
            if (BLOCK) __syncthreads(); else __syncwarp();
        }
    }
}

// Ascending bitonic sort of 32 * NPL keys held NPL per lane (element i = lane + 32 * j lives in key[j] of `lane`): strides below
// 32 exchange by shuffle, larger strides inside the lane.
template <int NPL>
__device__ __forceinline__ void warp_bitonic_sort_regs(uint64_t (&key)[NPL], uint32_t lane) {
#pragma unroll
    for (int size = 2; size <= 32 * NPL; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int js = stride >> 5;
#pragma unroll
                for (int j = 0; j < NPL; j++) {
                    if ((j & js) == 0) {
                        const bool up = ((static_cast<int>(lane) + 32 * j) & size) == 0;
                        const uint64_t a = key[j], b = key[j | js];
                        if ((a > b) == up) { key[j] = b; key[j | js] = a; }
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < NPL; j++) {
                    const bool up = ((static_cast<int>(lane) + 32 * j) & size) == 0;
                    const uint64_t other = __shfl_xor_sync(0xFFFFFFFFu, key[j], stride);
                    const bool lower = (lane & static_cast<uint32_t>(stride)) == 0;
                    key[j] = (lower == up) ? (key[j] < other ? key[j] : other) : (key[j] < other ? other : key[j]);
                }
            }
        }
    }
}

struct WarpSelect {
    uint64_t* buf;   // shared: [nsort] keys; [0,k) = current best (sorted after a flush), [k, k+cnt) = queue
    uint32_t nsort;  // power of two, >= k + 64
    uint32_t k;
    uint32_t cnt;    // warp-uniform
    uint64_t tau;    // warp-uniform: current k-th best key (KEY_SENTINEL until k keys were seen)

    __device__ __forceinline__ void init(uint64_t* b, uint32_t ns, uint32_t kk) {
        buf = b; nsort = ns; k = kk; cnt = 0; tau = KEY_SENTINEL;
        for (uint32_t i = lane_id(); i < kk; i += 32) buf[i] = KEY_SENTINEL;
        __syncwarp();
    }
    __device__ __forceinline__ void flush() {
        __syncwarp();
        for (uint32_t i = k + cnt + lane_id(); i < nsort; i += 32) buf[i] = KEY_SENTINEL;
        __syncwarp();
        bitonic_sort_keys<false>(buf, nsort, lane_id(), 32);
        tau = buf[k - 1];
        cnt = 0;
        __syncwarp();
    }
    // All 32 lanes must call (lanes without a candidate pass valid = false).
    __device__ __forceinline__ void offer(uint64_t key, bool valid) {
        bool hit = valid && (key < tau);
        uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
        if (m) {
            if (hit) buf[k + cnt + __popc(m & ((1u << lane_id()) - 1u))] = key;
            cnt += __popc(m);
            if (cnt + 32 > nsort - k) flush();
        }
    }
    static __host__ __device__ __forceinline__ uint32_t sort_size(uint32_t k) { return next_pow2(k + 64); }
};

}  // namespace annb
