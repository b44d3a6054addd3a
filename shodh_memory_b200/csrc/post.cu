// post.cu -- RetrievalEngine::search_ids post-processing on the device.
//
// Reference src/memory/retrieval.rs:927-961: for every (vector_id, distance) of the scan,
// similarity = -distance; map vector id -> MemoryId (a memory may own several chunk vectors); keep the
// HIGHEST similarity per memory; sort (similarity desc by total_cmp, MemoryId asc); truncate `limit`.
// One CTA per query; two in-smem bitonic sorts over <= 1024 entries with integer composite keys:
//   sort 1 by (MemoryId, distance key asc)  -> the first entry of each MemoryId run is its best chunk
//   sort 2 of those heads by (distance key asc, MemoryId asc)  == (similarity desc, MemoryId asc).
// similarity desc under total_cmp is exactly distance asc under total_cmp (negation flips the order),
// so the scan's order_key32(dist) is reused; the emitted similarity is the bit-exact negation.
#include "sg_internal.h"

namespace sg {
namespace {

struct Ent {
    uint64_t hi, lo;   // MemoryId (Uuid as two big-endian halves); all-ones = no mapping / padding
    uint32_t dkey;     // order_key32(distance)
    uint32_t pad;
};
SG_DEVICE_INLINE bool lt_mem_then_key(const Ent &a, const Ent &b)
{
    if (a.hi != b.hi) return a.hi < b.hi;
    if (a.lo != b.lo) return a.lo < b.lo;
    return a.dkey < b.dkey;
}
SG_DEVICE_INLINE bool lt_key_then_mem(const Ent &a, const Ent &b)
{
    if (a.dkey != b.dkey) return a.dkey < b.dkey;
    if (a.hi != b.hi) return a.hi < b.hi;
    return a.lo < b.lo;
}
template <typename Less>
SG_DEVICE_INLINE void block_sort(Ent *buf, uint32_t n, Less less)
{
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
                const uint32_t pos = 2 * i - (i & (stride - 1));
                const Ent a = buf[pos], b = buf[pos + stride];
                const bool up = (pos & size) == 0;
                if (less(b, a) == up) {
                    buf[pos] = b;
                    buf[pos + stride] = a;
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(512)
search_ids_post_kernel(const uint32_t *__restrict__ ids, const float *__restrict__ dist, const uint32_t *__restrict__ counts,
                       uint32_t k, const uint64_t *__restrict__ vec2mem, uint64_t n_map, uint32_t id_base, uint32_t limit,
                       uint64_t *__restrict__ mem_out, float *__restrict__ sim_out, uint32_t *__restrict__ cnt_out)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Ent *e = reinterpret_cast<Ent *>(smem_raw);
    __shared__ uint32_t s_heads;
    const uint32_t q = blockIdx.x, n = counts[q] < k ? counts[q] : k;
    const uint32_t M = next_pow2(k < 2 ? 2 : k);
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
        Ent x;
        x.hi = x.lo = ~0ull;
        x.dkey = 0xFFFFFFFFu;
        x.pad = 0;
        if (i < n) {
            const uint64_t vid = (uint64_t)ids[(size_t)q * k + i] - id_base;
            if (vid < n_map) {
                x.hi = vec2mem[2 * vid];
                x.lo = vec2mem[2 * vid + 1];
                x.dkey = order_key32(dist[(size_t)q * k + i]);
                if (x.hi == ~0ull && x.lo == ~0ull) x.dkey = 0xFFFFFFFFu;   // get_memory_id -> None
            }
        }
        e[i] = x;
    }
    if (threadIdx.x == 0) s_heads = 0;
    __syncthreads();
    block_sort(e, M, lt_mem_then_key);
    // heads of the MemoryId runs keep their key; everything else is pushed to the end
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
        const Ent x = e[i];
        const bool mapped = !(x.hi == ~0ull && x.lo == ~0ull);
        bool head = mapped && (i == 0 || e[i - 1].hi != x.hi || e[i - 1].lo != x.lo);
        if (head) atomicAdd(&s_heads, 1u);
        e[i].pad = head ? 1u : 0u;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x)
        if (!e[i].pad) {
            e[i].hi = e[i].lo = ~0ull;
            e[i].dkey = 0xFFFFFFFFu;
        }
    __syncthreads();
    block_sort(e, M, lt_key_then_mem);
    const uint32_t heads = s_heads, r = heads < limit ? heads : limit;
    for (uint32_t i = threadIdx.x; i < limit; i += blockDim.x) {
        const bool ok = i < r;
        mem_out[((size_t)q * limit + i) * 2] = ok ? e[i].hi : ~0ull;
        mem_out[((size_t)q * limit + i) * 2 + 1] = ok ? e[i].lo : ~0ull;
        sim_out[(size_t)q * limit + i] = ok ? -key32_to_dist(e[i].dkey) : 0.0f;
    }
    if (threadIdx.x == 0) cnt_out[q] = r;
}


// ---- RRFusion::fuse (src/memory/hybrid_search.rs:557-594): the Layer-3 -> Layer-4 hand-off ----------
// Ranked lists (BM25 leg from the host, vector leg from search_ids) -> union of MemoryIds with
// score = sum over lists, in list order, of w_l / (k + (rank+1) as f32)   (f32, round-to-nearest, no FMA),
// sorted (score total_cmp desc, MemoryId asc).  One CTA; <= 2048 entries:
//   sort 1 by (MemoryId, sequence)   -> each id's contributions become a run in accumulation order
//   run heads add their run left to right starting from 0.0 (HashMap `or_insert(0.0) += c`)
//   sort 2 of the heads by (~order_key32(score), MemoryId).
// All-ones IS a legal MemoryId here (the reference's tie-break test uses it), so validity is a flag.
struct RrfEnt {
    uint64_t hi, lo;
    uint32_t a;        // sort 1: sequence number; sort 2: ~order_key32(score)
    uint32_t dead;     // 1 = padding / not a run head
    float c;           // sort 1: contribution; sort 2: fused score
    uint32_t pad;
};
SG_DEVICE_INLINE bool rrf_lt(const RrfEnt &x, const RrfEnt &y)
{
    if (x.dead != y.dead) return x.dead < y.dead;
    if (x.hi != y.hi) return x.hi < y.hi;
    if (x.lo != y.lo) return x.lo < y.lo;
    return x.a < y.a;
}
SG_DEVICE_INLINE bool rrf_lt_score(const RrfEnt &x, const RrfEnt &y)
{
    if (x.dead != y.dead) return x.dead < y.dead;
    if (x.a != y.a) return x.a < y.a;
    if (x.hi != y.hi) return x.hi < y.hi;
    return x.lo < y.lo;
}
template <typename Less>
SG_DEVICE_INLINE void rrf_sort(RrfEnt *buf, uint32_t n, Less less)
{
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
                const uint32_t pos = 2 * i - (i & (stride - 1));
                const RrfEnt a = buf[pos], b = buf[pos + stride];
                const bool up = (pos & size) == 0;
                if (less(b, a) == up) {
                    buf[pos] = b;
                    buf[pos + stride] = a;
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(512)
rrf_fuse_kernel(float k, const float *__restrict__ weights, uint32_t n_weights, const uint64_t *__restrict__ ids,
                const uint32_t *__restrict__ offsets, uint32_t n_lists, uint32_t M, uint64_t *__restrict__ ids_out,
                float *__restrict__ score_out, uint32_t *__restrict__ n_out)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    RrfEnt *e = reinterpret_cast<RrfEnt *>(smem_raw);
    __shared__ float s_w[64];
    __shared__ uint32_t s_heads;
    const uint32_t n_total = offsets[n_lists];
    if (threadIdx.x == 0) {
        // RRFusion::new (:537-551): left-fold sum, w/sum, or 1/len when the sum is not positive
        float sum = 0.0f;
        for (uint32_t i = 0; i < n_weights; ++i) sum = __fadd_rn(sum, weights[i]);
        for (uint32_t i = 0; i < n_weights; ++i)
            s_w[i] = sum > 0.0f ? __fdiv_rn(weights[i], sum) : __fdiv_rn(1.0f, (float)n_weights);
        s_heads = 0;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
        RrfEnt x;
        x.hi = x.lo = 0;
        x.a = i;
        x.dead = 1;
        x.c = 0.0f;
        x.pad = 0;
        if (i < n_total) {
            uint32_t l = 0;
            while (l + 1 < n_lists && i >= offsets[l + 1]) ++l;
            const uint32_t rank = i - offsets[l];
            const float w = l < n_weights ? s_w[l] : 1.0f;                     // weights.get(idx).unwrap_or(1.0)
            x.c = __fdiv_rn(w, __fadd_rn(k, (float)(rank + 1)));
            x.hi = ids[2 * (size_t)i];
            x.lo = ids[2 * (size_t)i + 1];
            x.dead = 0;
        }
        e[i] = x;
    }
    __syncthreads();
    rrf_sort(e, M, rrf_lt);
    // run heads accumulate their run in sequence order
    float my_sum[4];
    bool my_head[4];
#pragma unroll
    for (uint32_t r = 0; r < 4; ++r) {
        const uint32_t i = threadIdx.x + r * blockDim.x;
        my_head[r] = false;
        my_sum[r] = 0.0f;
        if (i < M && !e[i].dead) {
            const RrfEnt x = e[i];
            my_head[r] = i == 0 || e[i - 1].hi != x.hi || e[i - 1].lo != x.lo;
            if (my_head[r]) {
                float sum = 0.0f;
                for (uint32_t j = i; j < M && !e[j].dead && e[j].hi == x.hi && e[j].lo == x.lo; ++j) sum = __fadd_rn(sum, e[j].c);
                my_sum[r] = sum;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (uint32_t r = 0; r < 4; ++r) {
        const uint32_t i = threadIdx.x + r * blockDim.x;
        if (i < M) {
            if (my_head[r]) {
                e[i].c = my_sum[r];
                e[i].a = ~order_key32(my_sum[r]);
                atomicAdd(&s_heads, 1u);
            } else {
                e[i].dead = 1;
            }
        }
    }
    __syncthreads();
    rrf_sort(e, M, rrf_lt_score);
    const uint32_t heads = s_heads;
    for (uint32_t i = threadIdx.x; i < heads; i += blockDim.x) {
        ids_out[2 * (size_t)i] = e[i].hi;
        ids_out[2 * (size_t)i + 1] = e[i].lo;
        score_out[i] = e[i].c;
    }
    if (threadIdx.x == 0) *n_out = heads;
}

}  // namespace

cudaError_t launch_search_ids_post(const uint32_t *d_ids, const float *d_dist, const uint32_t *d_counts, uint32_t b, uint32_t k,
                                   const uint64_t *d_vec2mem, uint64_t n_map, uint32_t id_base, uint32_t limit,
                                   uint64_t *d_mem_out, float *d_sim_out, uint32_t *d_cnt_out, cudaStream_t st)
{
    if (b == 0) return cudaSuccess;
    uint32_t M = 2;
    while (M < k) M <<= 1;
    const size_t smem = (size_t)M * sizeof(Ent);
    const uint32_t threads = M / 2 < 32 ? 32 : (M / 2 > 512 ? 512 : M / 2);
    search_ids_post_kernel<<<b, threads, smem, st>>>(d_ids, d_dist, d_counts, k, d_vec2mem, n_map, id_base, limit, d_mem_out,
                                                     d_sim_out, d_cnt_out);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_rrf_fuse(float k, const float *d_weights, uint32_t n_weights, const uint64_t *d_ids, const uint32_t *d_offsets,
                            uint32_t n_lists, uint32_t n_total, uint64_t *d_ids_out, float *d_score_out, uint32_t *d_n_out,
                            cudaStream_t st)
{
    uint32_t M = 2;
    while (M < n_total) M <<= 1;                       // n_total <= 2048 (checked by the caller)
    const size_t smem = (size_t)M * sizeof(RrfEnt);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(rrf_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const uint32_t threads = M / 4 < 32 ? 32 : (M / 4 > 512 ? 512 : M / 4);   // each thread owns <= 4 entries
    rrf_fuse_kernel<<<1, threads, smem, st>>>(k, d_weights, n_weights, d_ids, d_offsets, n_lists, M, d_ids_out, d_score_out, d_n_out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace sg
