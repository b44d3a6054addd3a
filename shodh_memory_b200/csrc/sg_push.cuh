// sg_push.cuh -- device side of the per-shard top-k exchange over NVLink peer memory.
//
// Symmetric buffer layout per rank, in 32-bit words, for one (b, k):
//   data [2 parities][W sources][b][2k]   (k ids, then k distance bit patterns)
//   flags[2 parities][W sources][b]       (epoch of the data; 0 = never written)
// A kernel that has just written a query's final (ids, dist)[k] calls push_query(): the CTA copies the k results
// into every rank's buffer with peer stores, fences at system scope and publishes the (source rank, query) flag
// with a release store.  exchange_wait_merge_kernel (exchange.cu) on every rank spins on the W flags of a query
// and merges the W lists.  Buffers are double-buffered by epoch parity: no barrier between consecutive searches.
#pragma once
#include "sg_internal.h"

namespace sg {

SG_DEVICE_INLINE size_t xchg_data_off(uint32_t parity, uint32_t src, uint32_t q, uint32_t W, uint32_t b, uint32_t k)
{
    return (((size_t)parity * W + src) * b + q) * 2 * k;
}
SG_DEVICE_INLINE size_t xchg_flag_off(uint32_t parity, uint32_t src, uint32_t q, uint32_t W, uint32_t b, uint32_t k)
{
    return (size_t)2 * W * b * 2 * k + ((size_t)parity * W + src) * b + q;
}
SG_DEVICE_INLINE void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
SG_DEVICE_INLINE uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Called by EVERY thread of the CTA after the CTA wrote ids_row[0..k) / dist_row[0..k) (global memory) for query q.
SG_DEVICE_INLINE void push_query(const PushArgs &p, uint32_t q, const uint32_t *ids_row, const float *dist_row)
{
    if (p.world == 0) return;
    __syncthreads();                                   // the rows were written by other threads of this CTA
    const uint32_t W = p.world, k = p.k, parity = p.epoch & 1u;
    q += p.q0;
    for (uint32_t i = threadIdx.x; i < 2 * k * W; i += blockDim.x) {
        const uint32_t r = i / (2 * k), e = i - r * 2 * k;
        const uint32_t v = e < k ? ids_row[e] : __float_as_uint(dist_row[e - k]);
        p.peer[r][xchg_data_off(parity, p.rank, q, W, p.b, k) + e] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < W)
        st_release_sys(p.peer[threadIdx.x] + xchg_flag_off(parity, p.rank, q, W, p.b, k), p.epoch);
}

}  // namespace sg
