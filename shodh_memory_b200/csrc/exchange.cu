// exchange.cu -- per-shard top-k exchange + k-way merge over NVLink peer memory.
//
// The multi-GPU recall step is "local sweep -> gather every shard's (id, dist)[b,k] -> merge"
// (SURVEY.md 8e).  Instead of an NCCL all-gather followed by a merge kernel, the ranks write their
// lists straight into each other's receive buffers with peer stores (buffers come from torch
// symmetric memory: cudaMalloc'd, mapped into every rank), publish a per-(rank, query) flag with a
// system-scope release store, and a merge CTA per query spins on the flags the peers publish into
// ITS buffer and merges the W lists as soon as they are all there -- transfer, synchronisation and
// merge pipeline query by query, and there is no collective launch on the critical path.
// Two forms:
//   * sg_store_search_sharded_device: the kernels that PRODUCE a query's final local result
//     (finalize_tc_kernel, merge_kernel) push it themselves (sg_push.cuh), so the exchange costs one
//     extra launch: exchange_wait_merge_kernel;
//   * sg_exchange_merge_topk_device: push + wait + merge of results that already sit in device
//     memory, in one kernel (used by the sharded graph search).
// Buffers are double-buffered by epoch parity, so no barrier is needed between consecutive searches.
// Layout: sg_push.cuh.
#include "sg_store_internal.h"
#include "sg_push.cuh"

namespace sg {
namespace {

struct XArgs {
    uint32_t *peer[8];        // base of every rank's symmetric buffer (peer[rank] is local)
    const uint32_t *ids;      // local results [b,k] (fused form only)
    const float *dist;
    uint32_t b, k, rank, world, epoch;
    uint32_t *ids_out; float *dist_out; uint32_t *counts_out;
};

// One CTA per query: wait until all W sources have published their list for it, then merge the W sorted
// lists (integer keys, the same (dist total_cmp, id) order as everywhere).
SG_DEVICE_INLINE void wait_and_merge(const XArgs &a, uint32_t q, uint64_t *keys, uint32_t *s_cnt)
{
    const uint32_t W = a.world, b = a.b, k = a.k, parity = a.epoch & 1u;
    const uint32_t *local = a.peer[a.rank];
    const uint32_t M = next_pow2(W * k < 2 ? 2 : W * k);
    if (threadIdx.x < W) {
        const uint32_t *f = local + xchg_flag_off(parity, threadIdx.x, q, W, b, k);
        while (ld_acquire_sys(f) != a.epoch) {
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
        uint64_t key = KEY_SENTINEL;
        if (i < W * k) {
            const uint32_t r = i / k, e = i - r * k;
            const uint32_t *src = local + xchg_data_off(parity, r, q, W, b, k);
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
    if (threadIdx.x == 0) *s_cnt = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) valid += __shfl_down_sync(0xFFFFFFFFu, valid, o);
    if ((threadIdx.x & 31) == 0 && valid) atomicAdd(s_cnt, valid);
    __syncthreads();
    if (threadIdx.x == 0) a.counts_out[q] = *s_cnt;
    __syncthreads();
}

// wait-only form: the pushes were done by the kernels that produced the results
__global__ void __launch_bounds__(256)
exchange_wait_merge_kernel(const __grid_constant__ XArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);
    __shared__ uint32_t s_cnt;
    pdl_wait();                                  // PDL launch: outputs of earlier searches on this stream are settled
    for (uint32_t q = blockIdx.x; q < a.b; q += gridDim.x) wait_and_merge(a, q, keys, &s_cnt);
}

// fused form.  grid <= #SMs (all CTAs co-resident: a CTA only ever waits for data that peers push BEFORE they wait)
__global__ void __launch_bounds__(256)
exchange_merge_kernel(const __grid_constant__ XArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);
    __shared__ uint32_t s_cnt;
    PushArgs p;
    for (uint32_t r = 0; r < 8; ++r) p.peer[r] = a.peer[r];
    p.rank = a.rank; p.world = a.world; p.epoch = a.epoch; p.b = a.b; p.k = a.k; p.q0 = 0;
    pdl_wait();                                  // PDL launch: the local results this kernel pushes are complete from here on
    for (uint32_t q = blockIdx.x; q < a.b; q += gridDim.x) {
        push_query(p, q, a.ids + (size_t)q * a.k, a.dist + (size_t)q * a.k);   // peer stores over NVLink + flag
        wait_and_merge(a, q, keys, &s_cnt);
    }
}

int fill_xargs(XArgs &a, const uint64_t *peer_buffers, uint32_t rank, uint32_t world, uint32_t epoch, uint32_t b, uint32_t k,
               uint32_t *d_ids_out, float *d_dist_out, uint32_t *d_counts_out)
{
    if (!peer_buffers || world == 0 || world > 8 || rank >= world || epoch == 0) {
        set_error("exchange: bad arguments (world 1..8, rank < world, epoch >= 1)");
        return SG_ERR_INVALID;
    }
    if (b && (!d_ids_out || !d_dist_out || !d_counts_out)) return SG_ERR_INVALID;
    if (k == 0 || k > 1024) {
        set_error("k=%u outside the supported range 1..1024", k);
        return SG_ERR_UNSUPPORTED;
    }
    a = XArgs{};
    for (uint32_t r = 0; r < world; ++r) a.peer[r] = reinterpret_cast<uint32_t *>(peer_buffers[r]);
    a.b = b; a.k = k; a.rank = rank; a.world = world; a.epoch = epoch;
    a.ids_out = d_ids_out; a.dist_out = d_dist_out; a.counts_out = d_counts_out;
    return SG_OK;
}

template <typename K>
int launch_x(K kern, const XArgs &a, int device, cudaStream_t st)
{
    uint32_t M = 2;
    while (M < a.world * a.k) M <<= 1;
    const size_t smem = (size_t)M * 8;
    if (smem > 48 * 1024) SG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0;
    int rc = check_device(device, &sms);
    if (rc != SG_OK) return rc;
    const uint32_t grid = a.b < (uint32_t)sms ? a.b : (uint32_t)sms;     // co-resident
    SG_CUDA_TRY(launch_pdl(kern, dim3(grid), dim3(256), smem, st, a));
    count_launch();
    return SG_OK;
}

}  // namespace

// search.cu-side helper: run the wait+merge step after a search that pushed its own results
int exchange_wait_merge(int device, const PushArgs &p, uint32_t *d_ids_out, float *d_dist_out, uint32_t *d_counts_out,
                        cudaStream_t st)
{
    XArgs a{};
    for (uint32_t r = 0; r < p.world; ++r) a.peer[r] = p.peer[r];
    a.b = p.b; a.k = p.k; a.rank = p.rank; a.world = p.world; a.epoch = p.epoch;
    a.ids_out = d_ids_out; a.dist_out = d_dist_out; a.counts_out = d_counts_out;
    return launch_x(exchange_wait_merge_kernel, a, device, st);
}

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
    XArgs a;
    int rc = fill_xargs(a, peer_buffers, rank, world, epoch, b, k, d_ids_out, d_dist_out, d_counts_out);
    if (rc != SG_OK) return rc;
    if (b && (!d_ids || !d_dist)) return SG_ERR_INVALID;
    if (b == 0) return SG_OK;
    SG_CUDA_TRY(cudaSetDevice(device));
    a.ids = d_ids;
    a.dist = d_dist;
    return launch_x(exchange_merge_kernel, a, device, static_cast<cudaStream_t>(cuda_stream));
}

}  // extern "C"
