// exchange.cu -- fused per-shard top-k exchange + k-way merge over NVLink peer memory.
//
// The multi-GPU recall step is "local sweep -> gather every shard's (id, dist)[b,k] -> merge"
// (SURVEY.md 8e).  Instead of an NCCL all-gather followed by a merge kernel, ONE kernel per rank
// does both: each CTA takes a query, PUSHES this rank's k results for it straight into every peer's
// receive buffer with peer stores (buffers come from torch symmetric memory: cudaMalloc'd, mapped
// into every rank), publishes a per-(rank, query) flag with a system-scope release store, spins on
// the flags the peers publish into ITS buffer, and merges the W lists as soon as they are all
// there -- transfer, synchronisation and merge pipeline query by query, and there is no collective
// launch on the critical path.  Buffers are double-buffered by epoch parity, so no barrier is
// needed between consecutive searches.
//
// Symmetric buffer layout per rank, in 32-bit words, for one (b, k):
//   data [2 parities][W sources][b][2k]   (k ids, then k distance bit patterns)
//   flags[2 parities][W sources][b]       (epoch of the data; 0 = never written)
#include "sg_internal.h"

namespace sg {
namespace {

struct XArgs {
    uint32_t *peer[8];        // base of every rank's symmetric buffer (peer[rank] is local)
    const uint32_t *ids;      // local results [b,k]
    const float *dist;
    uint32_t b, k, rank, world, epoch;
    uint32_t *ids_out; float *dist_out; uint32_t *counts_out;
};

SG_DEVICE_INLINE size_t data_off(uint32_t parity, uint32_t src, uint32_t q, uint32_t W, uint32_t b, uint32_t k)
{
    return (((size_t)parity * W + src) * b + q) * 2 * k;
}
SG_DEVICE_INLINE size_t flag_off(uint32_t parity, uint32_t src, uint32_t q, uint32_t W, uint32_t b, uint32_t k)
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

// grid <= #SMs (all CTAs co-resident: a CTA only ever waits for data that peers push BEFORE they wait)
__global__ void __launch_bounds__(256)
exchange_merge_kernel(const XArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);
    const uint32_t W = a.world, b = a.b, k = a.k, parity = a.epoch & 1u;
    uint32_t *local = a.peer[a.rank];
    const uint32_t M = next_pow2(W * k < 2 ? 2 : W * k);
    for (uint32_t q = blockIdx.x; q < b; q += gridDim.x) {
        // (1) push my list for query q into every rank's buffer (peer stores over NVLink)
        for (uint32_t i = threadIdx.x; i < 2 * k * W; i += blockDim.x) {
            const uint32_t r = i / (2 * k), e = i - r * 2 * k;
            const uint32_t v = e < k ? a.ids[(size_t)q * k + e] : __float_as_uint(a.dist[(size_t)q * k + e - k]);
            a.peer[r][data_off(parity, a.rank, q, W, b, k) + e] = v;
        }
        __threadfence_system();
        __syncthreads();
        // (2) publish, (3) wait for every source
        if (threadIdx.x < W) {
            st_release_sys(a.peer[threadIdx.x] + flag_off(parity, a.rank, q, W, b, k), a.epoch);
            const uint32_t *f = local + flag_off(parity, threadIdx.x, q, W, b, k);
            while (ld_acquire_sys(f) != a.epoch) {
            }
        }
        __syncthreads();
        // (4) merge the W sorted lists: integer keys, same (dist total_cmp, id) order as everywhere
        for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
            uint64_t key = KEY_SENTINEL;
            if (i < W * k) {
                const uint32_t r = i / k, e = i - r * k;
                const uint32_t *src = local + data_off(parity, r, q, W, b, k);
                const uint32_t id = __ldcv(src + e);
                if (id != 0xFFFFFFFFu) key = make_key(order_key32(__uint_as_float(__ldcv(src + k + e))), id);
            }
            keys[i] = key;
        }
        __syncthreads();
        block_bitonic_sort(keys, M);
        uint32_t valid = 0;
        for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
            const uint64_t key = keys[i];
            const bool ok = key != KEY_SENTINEL;
            a.ids_out[(size_t)q * k + i] = ok ? (uint32_t)(key & 0xFFFFFFFFu) : 0xFFFFFFFFu;
            a.dist_out[(size_t)q * k + i] = ok ? key32_to_dist((uint32_t)(key >> 32)) : __uint_as_float(0x7F800000u);
            valid += ok;
        }
        __shared__ uint32_t s_cnt;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        for (int o = 16; o > 0; o >>= 1) valid += __shfl_down_sync(0xFFFFFFFFu, valid, o);
        if ((threadIdx.x & 31) == 0 && valid) atomicAdd(&s_cnt, valid);
        __syncthreads();
        if (threadIdx.x == 0) a.counts_out[q] = s_cnt;
        __syncthreads();
    }
}

}  // namespace
}  // namespace sg

using namespace sg;

extern "C" {

uint64_t sg_exchange_buffer_words(uint32_t world, uint32_t b, uint32_t k)
{
    return (uint64_t)2 * world * b * 2 * k + (uint64_t)2 * world * b;
}

int sg_exchange_merge_topk_device(int device, const uint64_t *peer_buffers, uint32_t rank, uint32_t world, uint32_t epoch,
                                  const uint32_t *d_ids, const float *d_dist, uint32_t b, uint32_t k, uint32_t *d_ids_out,
                                  float *d_dist_out, uint32_t *d_counts_out, void *cuda_stream)
{
    if (!peer_buffers || world == 0 || world > 8 || rank >= world || epoch == 0) {
        set_error("sg_exchange_merge_topk_device: bad arguments (world 1..8, rank < world, epoch >= 1)");
        return SG_ERR_INVALID;
    }
    if (b && (!d_ids || !d_dist || !d_ids_out || !d_dist_out || !d_counts_out)) return SG_ERR_INVALID;
    if (k == 0 || k > 1024) {
        set_error("k=%u outside the supported range 1..1024", k);
        return SG_ERR_UNSUPPORTED;
    }
    if (b == 0) return SG_OK;
    SG_CUDA_TRY(cudaSetDevice(device));
    XArgs a{};
    for (uint32_t r = 0; r < world; ++r) a.peer[r] = reinterpret_cast<uint32_t *>(peer_buffers[r]);
    a.ids = d_ids; a.dist = d_dist; a.b = b; a.k = k; a.rank = rank; a.world = world; a.epoch = epoch;
    a.ids_out = d_ids_out; a.dist_out = d_dist_out; a.counts_out = d_counts_out;
    uint32_t M = 2;
    while (M < world * k) M <<= 1;
    const size_t smem = (size_t)M * 8;
    if (smem > 48 * 1024)
        SG_CUDA_TRY(cudaFuncSetAttribute(exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0;
    SG_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const uint32_t grid = b < (uint32_t)sms ? b : (uint32_t)sms;
    exchange_merge_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(cuda_stream)>>>(a);
    count_launch();
    SG_CUDA_TRY(cudaGetLastError());
    return SG_OK;
}

}  // extern "C"
