// sg_common.cuh -- shared device helpers for libshodh_gpu (sm_100a only).
// PTX wrappers for mbarrier / TMA, order-preserving keys, small in-smem sorts.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>

#include "../../include/shodh_gpu.h"

#ifndef SG_DEVICE_INLINE
#define SG_DEVICE_INLINE __device__ __forceinline__
#endif

namespace sg {

// ----------------------------------------------------------------------------------------
// error plumbing (host)
// ----------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(unsigned n = 1);

#define SG_CUDA_TRY(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            sg::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,              \
                          cudaGetErrorString(_e));                                          \
            return (_e == cudaErrorMemoryAllocation) ? SG_ERR_OOM : SG_ERR_CUDA;            \
        }                                                                                   \
    } while (0)

// ----------------------------------------------------------------------------------------
// programmatic dependent launch: the small kernels behind a sweep (re-score, flagged re-run, merges, exchange) are
// launched with programmatic stream serialization, so their launch latency overlaps the kernel before them; each
// calls pdl_wait() before it touches anything its predecessor wrote (a no-op under a normal launch).
// ----------------------------------------------------------------------------------------
SG_DEVICE_INLINE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// lets the NEXT kernel of the stream start its prologue now (it still blocks in its own pdl_wait() until this grid has
// completed and flushed); fires once every CTA of this grid has called it or exited
SG_DEVICE_INLINE void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename K, typename... Args>
inline cudaError_t launch_pdl(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static const bool no_pdl = getenv("SG_NO_PDL") != nullptr;      // measurement switch: plain stream-ordered launches
    cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

// ----------------------------------------------------------------------------------------
// keys: (distance, id) -> u64 whose unsigned order is (dist total_cmp asc, id asc)
// (src/vector_db/vamana.rs:1184, :1664-1673)
// ----------------------------------------------------------------------------------------
constexpr uint64_t KEY_SENTINEL = 0xFFFFFFFFFFFFFFFFull;  // sorts after every real candidate

SG_DEVICE_INLINE uint32_t order_key32(float dist)
{
    uint32_t b = __float_as_uint(dist);
    // IEEE totalOrder as unsigned order: negative -> ~b, non-negative -> b | sign
    return b ^ (uint32_t)(((int32_t)b >> 31) | 0x80000000);
}
SG_DEVICE_INLINE float key32_to_dist(uint32_t k)
{
    uint32_t b = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
    return __uint_as_float(b);
}
SG_DEVICE_INLINE uint64_t make_key(uint32_t key32, uint32_t id) { return ((uint64_t)key32 << 32) | id; }

// -dot with NaN canonicalised to +NaN (sorts last; x86 default-NaN dot gives the same order)
SG_DEVICE_INLINE float neg_dot_to_dist(float dot)
{
    return (dot != dot) ? __uint_as_float(0x7FC00000u) : -dot;
}

// ----------------------------------------------------------------------------------------
// shared-memory address / mbarrier / TMA (sm_90+ PTX, compiled for sm_100a)
// ----------------------------------------------------------------------------------------
SG_DEVICE_INLINE uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

SG_DEVICE_INLINE void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
SG_DEVICE_INLINE void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
SG_DEVICE_INLINE void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
SG_DEVICE_INLINE void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
SG_DEVICE_INLINE bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
SG_DEVICE_INLINE void mbar_wait(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
SG_DEVICE_INLINE void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 2-D tiled TMA load global -> shared, completion on an mbarrier (SASS: UTMALDG)
SG_DEVICE_INLINE void tma_load_2d(uint32_t dst_smem, const CUtensorMap *tmap, uint32_t bar, int32_t x,
                                  int32_t y)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst_smem), "l"((uint64_t)tmap), "r"(bar), "r"(x), "r"(y)
        : "memory");
}
// 2-D tiled TMA store shared -> global (bulk async-group completion; SASS: UTMASTG)
SG_DEVICE_INLINE void tma_store_2d(const CUtensorMap *tmap, uint32_t src_smem, int32_t x, int32_t y)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"((uint64_t)tmap), "r"(src_smem), "r"(x), "r"(y)
                 : "memory");
}
SG_DEVICE_INLINE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
SG_DEVICE_INLINE void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
SG_DEVICE_INLINE void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
SG_DEVICE_INLINE void tma_prefetch_desc(const CUtensorMap *tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tmap) : "memory");
}

// ----------------------------------------------------------------------------------------
// bitonic sorts of u64 keys in shared memory (ascending). n must be a power of two.
// ----------------------------------------------------------------------------------------
// one warp, warp-synchronous
SG_DEVICE_INLINE void warp_bitonic_sort(uint64_t *buf, uint32_t n, uint32_t lane)
{
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = lane; i < (n >> 1); i += 32) {
                uint32_t pos = 2 * i - (i & (stride - 1));
                uint64_t a = buf[pos], b = buf[pos + stride];
                bool up = (pos & size) == 0;
                if ((a > b) == up) {
                    buf[pos] = b;
                    buf[pos + stride] = a;
                }
            }
            __syncwarp();
        }
    }
}
// whole block (all threads must call)
SG_DEVICE_INLINE void block_bitonic_sort(uint64_t *buf, uint32_t n)
{
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
                uint32_t pos = 2 * i - (i & (stride - 1));
                uint64_t a = buf[pos], b = buf[pos + stride];
                bool up = (pos & size) == 0;
                if ((a > b) == up) {
                    buf[pos] = b;
                    buf[pos + stride] = a;
                }
            }
            __syncthreads();
        }
    }
}

__host__ __device__ __forceinline__ uint32_t next_pow2(uint32_t x)
{
    uint32_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

}  // namespace sg
