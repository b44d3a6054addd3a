// select.cu -- final top-k selection over per-CTA / per-shard candidate lists.
//
// Replaces `distances.sort_by(|a,b| a.1.total_cmp(&b.1).then_with(|| a.0.cmp(&b.0)));
// distances.truncate(k)` (reference src/vector_db/vamana.rs:1184-1185) for the
// survivors of the sweep, and serves as the k-way merge of per-shard results
// (SURVEY.md 8e).  All ordering is on 64-bit integer keys (order_key32(dist) << 32 | id),
// whose unsigned order IS (dist total_cmp asc, id asc): no float compare, no ties.
#include <mutex>

#include "sg_internal.h"
#include "sg_push.cuh"

namespace sg {
namespace {

constexpr uint32_t SORT_MAX = 8192;  // keys per in-smem sort (64 KB)

// One CTA per (query, group).  With n_groups == 1 the CTA reduces all of the query's keys and
// writes decoded (id, dist) results; with n_groups > 1 (first level of a two-level merge) it
// reduces keys [grp*group_keys, ...) of the query and writes its k best RAW keys to keys_out.
template <bool PAIRS>
__global__ void __launch_bounds__(1024, 1)
merge_kernel(const uint64_t *__restrict__ lists, const uint32_t *__restrict__ ids_in,
             const float *__restrict__ dist_in, uint32_t n_per_query_total, uint32_t group_keys, uint32_t n_groups,
             uint32_t k, uint32_t id_base, uint64_t list_stride, uint64_t query_stride,
             uint64_t *__restrict__ keys_out, uint32_t *__restrict__ ids_out, float *__restrict__ dist_out,
             uint32_t *__restrict__ counts, const uint32_t *__restrict__ slot_count,
             const uint32_t *__restrict__ slot_to_query, const __grid_constant__ PushArgs push)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t *buf = reinterpret_cast<uint64_t *>(smem_raw);
    pdl_wait();                                                   // PDL launch: the producers' lists are visible from here on
    if (slot_count && blockIdx.x >= *slot_count) return;          // flagged form: nothing to do for this slot
    const uint32_t q = blockIdx.x / n_groups, grp = blockIdx.x - q * n_groups;
    const uint32_t q_out = slot_to_query ? slot_to_query[q] : q;  // where the result goes
    const uint32_t first = grp * group_keys;
    const uint32_t n_per_query = first >= n_per_query_total ? 0
                                 : (n_per_query_total - first < group_keys ? n_per_query_total - first : group_keys);
    const size_t base = (size_t)q * n_per_query_total + first;

    uint32_t kept = 0, pos = 0;
    // Fast path for SORTED lists of k slots each (every producer in this library writes them sorted):
    // take the first c = ceil(k / n_lists) entries of every list; the k-th smallest of those, T0, has at
    // least k entries at or below it, so the global top-k all satisfy key <= T0.  Filter everything by T0
    // (typically ~k..2k survivors) and sort only the survivors.  Falls through to the chunked bitonic
    // path if the survivors do not fit (pathological inputs).
    const uint32_t n_lists_here = n_per_query / k;
    __shared__ uint32_t s_surv;
    __shared__ uint64_t s_T0;
    if (n_lists_here >= 2 && n_lists_here * k == n_per_query && n_per_query > 2 * k) {
        const uint32_t c = (k + n_lists_here - 1) / n_lists_here;
        const uint32_t nh = c * n_lists_here;                       // <= k + n_lists
        if (nh <= SORT_MAX) {
            auto load_key = [&](uint32_t e) -> uint64_t {
                if constexpr (PAIRS) {
                    const uint32_t ee = first + e, l = ee / k, j = ee - l * k;
                    const size_t at = (size_t)l * list_stride + (size_t)q * query_stride + j;
                    const uint32_t id = ids_in[at];
                    return (id == 0xFFFFFFFFu) ? KEY_SENTINEL : make_key(order_key32(dist_in[at]), id);
                } else {
                    return lists[base + e];
                }
            };
            const uint32_t MH = next_pow2(nh < 2 ? 2 : nh);
            for (uint32_t i = threadIdx.x; i < MH; i += blockDim.x)
                buf[i] = i < nh ? load_key((i / c) * k + (i % c)) : KEY_SENTINEL;
            if (threadIdx.x == 0) s_surv = 0;
            __syncthreads();
            block_bitonic_sort(buf, MH);
            if (threadIdx.x == 0) s_T0 = buf[k - 1];
            __syncthreads();
            const uint64_t T0 = s_T0;
            __syncthreads();
            for (uint32_t e = threadIdx.x; e < n_per_query; e += blockDim.x) {
                const uint64_t key = load_key(e);
                if (key != KEY_SENTINEL && key <= T0) {
                    const uint32_t slot = atomicAdd(&s_surv, 1u);
                    if (slot < SORT_MAX) buf[slot] = key;
                }
            }
            __syncthreads();
            const uint32_t ns = s_surv;
            if (ns <= SORT_MAX) {
                const uint32_t M = next_pow2(ns < 2 ? 2 : ns);
                for (uint32_t i = ns + threadIdx.x; i < M; i += blockDim.x) buf[i] = KEY_SENTINEL;
                __syncthreads();
                block_bitonic_sort(buf, M);
                kept = ns < k ? ns : k;
                pos = n_per_query;                                   // done: skip the chunked path
            }
            __syncthreads();
        }
    }
    while (pos < n_per_query) {
        uint32_t take = n_per_query - pos;
        if (take > SORT_MAX - kept) take = SORT_MAX - kept;
        for (uint32_t i = threadIdx.x; i < take; i += blockDim.x) {
            uint64_t key;
            if constexpr (PAIRS) {
                const uint32_t e = first + pos + i, l = e / k, j = e - l * k;
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
    if (n_groups > 1) {
        uint64_t *out = keys_out + ((size_t)q * n_groups + grp) * k;
        for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) out[i] = i < kept ? buf[i] : KEY_SENTINEL;
        return;
    }
    uint32_t local = 0;
    for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
        uint64_t key = i < kept ? buf[i] : KEY_SENTINEL;
        bool valid = key != KEY_SENTINEL;
        ids_out[(size_t)q_out * k + i] = valid ? (uint32_t)(key & 0xFFFFFFFFu) + id_base : 0xFFFFFFFFu;
        dist_out[(size_t)q_out * k + i] = valid ? key32_to_dist((uint32_t)(key >> 32)) : __uint_as_float(0x7F800000u);
        local += valid;
    }
    // block reduce the count of valid results
    __shared__ uint32_t s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xFFFFFFFFu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) counts[q_out] = s_cnt;
    // multi-GPU: this is the query's final local result -- hand it to the peers right away
    push_query(push, q_out, ids_out + (size_t)q_out * k, dist_out + (size_t)q_out * k);
}

template <bool PAIRS>
cudaError_t launch(const uint64_t *lists, const uint32_t *ids_in, const float *dist_in, uint32_t b,
                   uint32_t n_per_query, uint32_t group_keys, uint32_t n_groups, uint32_t k, uint32_t id_base,
                   uint64_t list_stride, uint64_t query_stride, uint64_t *keys_out, uint32_t *ids, float *dist,
                   uint32_t *counts, cudaStream_t st, const uint32_t *slot_count = nullptr,
                   const uint32_t *slot_to_query = nullptr, const PushArgs *push = nullptr)
{
    if (b == 0) return cudaSuccess;
    const uint32_t per_cta = n_groups > 1 ? group_keys : n_per_query;
    uint32_t need = per_cta < SORT_MAX ? next_pow2(per_cta < 2 ? 2 : per_cta) : SORT_MAX;   // survivors <= per_cta
    if (need < k) need = next_pow2(k);
    size_t smem = (size_t)need * 8;
    uint32_t threads = need / 2 < 1024 ? (need / 2 < 32 ? 32 : need / 2) : 1024;
    auto kern = merge_kernel<PAIRS>;
    {
        static DevAttr da;                       // process-wide table: two stores on two threads may get here together
        cudaError_t e = ensure_smem(kern, da, (size_t)SORT_MAX * 8);
        if (e != cudaSuccess) return e;
    }
    PushArgs pa{};
    if (push && n_groups == 1) pa = *push;       // only the stage that writes the final (ids, dist) pushes
    cudaError_t le = launch_pdl(kern, dim3(b * n_groups), dim3(threads), smem, st, lists, ids_in, dist_in, n_per_query, group_keys,
                                n_groups, k, id_base, list_stride, query_stride, keys_out, ids, dist, counts, slot_count,
                                slot_to_query, pa);
    count_launch();
    return le;
}

}  // namespace

// how many first-level groups a (n_lists x k)-key merge is split into (1 = single level)
uint32_t merge_groups(uint32_t n_lists, uint32_t k)
{
    // with the sorted-list threshold path one CTA per query handles every shape this library produces
    // (first ceil(k/n_lists) entries of each list <= 8192 keys); keep the two-level split only beyond that
    if ((uint64_t)((k + (n_lists ? n_lists : 1) - 1) / (n_lists ? n_lists : 1)) * n_lists <= 8192) return 1;
    if ((uint64_t)n_lists * k <= 4096) return 1;
    uint32_t lists_per_group = 2048 / k ? 2048 / k : 1;
    if (lists_per_group < 2) lists_per_group = 2;
    return (n_lists + lists_per_group - 1) / lists_per_group;
}

cudaError_t launch_merge_keys(const uint64_t *d_lists, uint32_t b, uint32_t n_lists, uint32_t k, uint32_t id_base,
                              uint64_t *d_scratch, uint32_t *d_ids, float *d_dist, uint32_t *d_counts,
                              cudaStream_t st, const PushArgs *push)
{
    const uint32_t groups = d_scratch ? merge_groups(n_lists, k) : 1;
    if (groups > 1) {
        const uint32_t lists_per_group = (n_lists + groups - 1) / groups;
        cudaError_t e = launch<false>(d_lists, nullptr, nullptr, b, n_lists * k, lists_per_group * k, groups, k, 0, 0, 0,
                                      d_scratch, nullptr, nullptr, nullptr, st);
        if (e != cudaSuccess) return e;
        return launch<false>(d_scratch, nullptr, nullptr, b, groups * k, groups * k, 1, k, id_base, 0, 0, nullptr,
                             d_ids, d_dist, d_counts, st, nullptr, nullptr, push);
    }
    return launch<false>(d_lists, nullptr, nullptr, b, n_lists * k, n_lists * k, 1, k, id_base, 0, 0, nullptr, d_ids,
                         d_dist, d_counts, st, nullptr, nullptr, push);
}

cudaError_t launch_merge_keys_flagged(const uint64_t *d_lists, uint32_t b, uint32_t n_lists, uint32_t k, uint32_t id_base,
                                      const uint32_t *d_slot_count, const uint32_t *d_slot_to_query, uint32_t *d_ids,
                                      float *d_dist, uint32_t *d_counts, cudaStream_t st, const PushArgs *push)
{
    return launch<false>(d_lists, nullptr, nullptr, b, n_lists * k, n_lists * k, 1, k, id_base, 0, 0, nullptr, d_ids,
                         d_dist, d_counts, st, d_slot_count, d_slot_to_query, push);
}

cudaError_t launch_merge_pairs(const uint32_t *d_ids_in, const float *d_dist_in, uint32_t b, uint32_t n_lists,
                               uint32_t k, uint64_t list_stride, uint64_t query_stride, uint32_t *d_ids,
                               float *d_dist, uint32_t *d_counts, cudaStream_t st)
{
    return launch<true>(nullptr, d_ids_in, d_dist_in, b, n_lists * k, n_lists * k, 1, k, 0, list_stride, query_stride,
                        nullptr, d_ids, d_dist, d_counts, st);
}

}  // namespace sg
