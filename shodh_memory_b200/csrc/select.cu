// select.cu -- final top-k selection over per-CTA / per-shard candidate lists.
//
// Replaces `distances.sort_by(|a,b| a.1.total_cmp(&b.1).then_with(|| a.0.cmp(&b.0)));
// distances.truncate(k)` (reference src/vector_db/vamana.rs:1184-1185) for the
// survivors of the sweep, and serves as the k-way merge of per-shard results
// (SURVEY.md 8e).  All ordering is on 64-bit integer keys (order_key32(dist) << 32 | id),
// whose unsigned order IS (dist total_cmp asc, id asc): no float compare, no ties.
#include "sg_internal.h"

namespace sg {
namespace {

constexpr uint32_t SORT_MAX = 8192;  // keys per in-smem sort (64 KB)

template <bool PAIRS>
__global__ void __launch_bounds__(1024, 1)
merge_kernel(const uint64_t *__restrict__ lists, const uint32_t *__restrict__ ids_in,
             const float *__restrict__ dist_in, uint32_t n_per_query, uint32_t k, uint32_t id_base,
             uint64_t list_stride, uint64_t query_stride,
             uint32_t *__restrict__ ids_out, float *__restrict__ dist_out, uint32_t *__restrict__ counts)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t *buf = reinterpret_cast<uint64_t *>(smem_raw);
    const uint32_t q = blockIdx.x;
    const size_t base = (size_t)q * n_per_query;

    uint32_t kept = 0, pos = 0;
    while (pos < n_per_query) {
        uint32_t take = n_per_query - pos;
        if (take > SORT_MAX - kept) take = SORT_MAX - kept;
        for (uint32_t i = threadIdx.x; i < take; i += blockDim.x) {
            uint64_t key;
            if constexpr (PAIRS) {
                const uint32_t e = pos + i, l = e / k, j = e - l * k;
                const size_t at = (size_t)l * list_stride + (size_t)q * query_stride + j;
                uint32_t id = ids_in[at];
                key = (id == 0xFFFFFFFFu) ? KEY_SENTINEL : make_key(order_key32(dist_in[at]), id);
            } else {
                key = lists[base + pos + i];
            }
            buf[kept + i] = key;
        }
        uint32_t tot = kept + take;
        uint32_t M = next_pow2(tot < 2 ? 2 : tot);
        for (uint32_t i = tot + threadIdx.x; i < M; i += blockDim.x) buf[i] = KEY_SENTINEL;
        __syncthreads();
        block_bitonic_sort(buf, M);
        kept = tot < k ? tot : k;
        pos += take;
    }
    if (n_per_query == 0) {
        for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) buf[i] = KEY_SENTINEL;
        __syncthreads();
    }
    uint32_t local = 0;
    for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
        uint64_t key = i < kept ? buf[i] : KEY_SENTINEL;
        bool valid = key != KEY_SENTINEL;
        ids_out[(size_t)q * k + i] = valid ? (uint32_t)(key & 0xFFFFFFFFu) + id_base : 0xFFFFFFFFu;
        dist_out[(size_t)q * k + i] = valid ? key32_to_dist((uint32_t)(key >> 32)) : __uint_as_float(0x7F800000u);
        local += valid;
    }
    // block reduce the count of valid results
    __shared__ uint32_t s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xFFFFFFFFu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) counts[q] = s_cnt;
}

template <bool PAIRS>
cudaError_t launch(const uint64_t *lists, const uint32_t *ids_in, const float *dist_in, uint32_t b,
                   uint32_t n_per_query, uint32_t k, uint32_t id_base, uint64_t list_stride, uint64_t query_stride,
                   uint32_t *ids, float *dist, uint32_t *counts, cudaStream_t st)
{
    if (b == 0) return cudaSuccess;
    uint32_t need = n_per_query < SORT_MAX ? next_pow2(n_per_query < 2 ? 2 : n_per_query) : SORT_MAX;
    if (need < k) need = next_pow2(k);
    size_t smem = (size_t)need * 8;
    uint32_t threads = need / 2 < 1024 ? (need / 2 < 32 ? 32 : need / 2) : 1024;
    auto kern = merge_kernel<PAIRS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SORT_MAX * 8));
    if (e != cudaSuccess) return e;
    kern<<<b, threads, smem, st>>>(lists, ids_in, dist_in, n_per_query, k, id_base, list_stride, query_stride, ids,
                                   dist, counts);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_merge_keys(const uint64_t *d_lists, uint32_t b, uint32_t n_per_query, uint32_t k,
                              uint32_t id_base, uint32_t *d_ids, float *d_dist, uint32_t *d_counts,
                              cudaStream_t st)
{
    return launch<false>(d_lists, nullptr, nullptr, b, n_per_query, k, id_base, 0, 0, d_ids, d_dist, d_counts, st);
}

cudaError_t launch_merge_pairs(const uint32_t *d_ids_in, const float *d_dist_in, uint32_t b, uint32_t n_lists,
                               uint32_t k, uint64_t list_stride, uint64_t query_stride, uint32_t *d_ids,
                               float *d_dist, uint32_t *d_counts, cudaStream_t st)
{
    return launch<true>(nullptr, d_ids_in, d_dist_in, b, n_lists * k, k, 0, list_stride, query_stride, d_ids, d_dist,
                        d_counts, st);
}

}  // namespace sg
