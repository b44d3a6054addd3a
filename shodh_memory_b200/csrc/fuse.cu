// fuse.cu -- relevance fusion + final ranking of the scan's survivors.
//
// Replaces the per-candidate math and the sort/threshold/truncate of
// RelevanceEngine::surface_relevant_inner (reference src/relevance.rs:805-918):
//   LearnedWeights::fuse_scores_full  :529-594   (7 sigmoid-calibrated signals)
//   calibrate_score                   :601-606   (1/(1+exp(-10(x-0.5))), non-finite -> 0)
//   apply_recency_boost               :1524-1547 (integer hours, linear decay, cap 1.0)
//   graph boost                       :877-881   (x1.15 capped at 1.0 when entity_score > 0)
//   sort (score desc, created_at desc, id asc) :905-910 ; >= 0.25 :914-915 ; truncate :918
// All arithmetic is f32 with explicit rounding (__fadd_rn/__fmul_rn/__fdiv_rn: no FMA
// contraction, like rustc).  exp/log2 are evaluated in f64 and rounded once to f32, which
// agrees with glibc's expf/log2f (what Rust's f32::exp/log2 call) except in the rare
// half-ulp cases; the parity tolerance for fused scores is therefore 1e-6, stated in
// tests/test_fuse_gpu.py.  The work is ~100 candidates per query: one CTA, latency-bound.
#include "sg_internal.h"

namespace sg {
namespace {

SG_DEVICE_INLINE float calibrate(float score)
{
    if (!isfinite(score)) return 0.0f;
    float u = __fmul_rn(-10.0f, __fsub_rn(score, 0.5f));
    float e = (float)exp((double)u);
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
}

SG_DEVICE_INLINE float fuse_one(const sg_fuse_weights &w, float sem, float ent, float tag, float imp,
                                float mom, uint32_t access, float graph)
{
    float cs = calibrate(sem), ce = calibrate(ent), ct = calibrate(tag), ci = calibrate(imp);
    float nm = __fdiv_rn(__fadd_rn(mom, 1.0f), 2.0f);
    float am;
    if (nm > 0.65f)
        am = fminf(__fmul_rn(nm, 1.5f), 1.0f);
    else if (nm < 0.40f)
        am = fmaxf(__fmul_rn(nm, 0.3f), 0.0f);
    else
        am = nm;
    float cm = calibrate(am);
    float as = 0.0f;
    if (access != 0) {
        float la = (float)log2((double)__fadd_rn((float)access, 1.0f));
        as = fminf(__fdiv_rn(la, 4.0f), 1.0f);
    }
    float ca = calibrate(as), cg = calibrate(graph);
    float r = __fadd_rn(__fmul_rn(w.semantic, cs), __fmul_rn(w.entity, ce));
    r = __fadd_rn(r, __fmul_rn(w.tag, ct));
    r = __fadd_rn(r, __fmul_rn(w.importance, ci));
    r = __fadd_rn(r, __fmul_rn(w.momentum, cm));
    r = __fadd_rn(r, __fmul_rn(w.access_count, ca));
    r = __fadd_rn(r, __fmul_rn(w.graph_strength, cg));
    return isfinite(r) ? r : 0.0f;
}

SG_DEVICE_INLINE float recency(float base, int64_t now_s, int64_t created_s, uint64_t boost_hours, float mult)
{
    if (boost_hours == 0) return base;
    int64_t hours = (now_s - created_s) / 3600;  // truncates toward zero == chrono num_hours()
    uint64_t age_hours = (uint64_t)hours;        // negative -> huge, like `as u64`
    if (age_hours <= boost_hours) {
        float decay = __fsub_rn(1.0f, __fdiv_rn((float)age_hours, (float)boost_hours));
        float boost = __fadd_rn(1.0f, __fmul_rn(__fsub_rn(mult, 1.0f), decay));
        return fminf(__fmul_rn(base, boost), 1.0f);
    }
    return base;
}

struct RankKey {
    uint64_t hi, lo;
};
SG_DEVICE_INLINE bool key_gt(const RankKey &a, const RankKey &b)
{
    return a.hi > b.hi || (a.hi == b.hi && a.lo > b.lo);
}

__global__ void __launch_bounds__(1024, 1)
fuse_rank_kernel(FuseDev f, uint32_t n, sg_fuse_weights w, sg_fuse_config c, int64_t now_s,
                 uint32_t *order_out, float *score_out, uint32_t *n_out, float *all_out)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    RankKey *keys = reinterpret_cast<RankKey *>(smem_raw);
    const uint32_t M = next_pow2(n < 2 ? 2 : n);
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
        RankKey key;
        key.hi = ~0ull;
        key.lo = ~0ull;
        if (i < n) {
            float fused = fuse_one(w, f.semantic[i], f.entity[i], f.tag[i], f.importance[i], f.momentum[i],
                                   f.access[i], f.graph[i]);
            float boosted = recency(fused, now_s, f.created[i], c.recency_boost_hours, c.recency_boost_multiplier);
            float fin = f.entity[i] > 0.0f ? fminf(__fmul_rn(boosted, c.graph_boost_multiplier), 1.0f) : boosted;
            if (all_out) all_out[i] = fin;
            // ascending 128-bit key == (score desc by total_cmp, created_at desc, position asc)
            uint32_t sk = ~order_key32(fin);
            uint64_t ck = ~((uint64_t)f.created[i] ^ 0x8000000000000000ull);
            key.hi = ((uint64_t)sk << 32) | (ck >> 32);
            key.lo = (ck << 32) | i;
        }
        keys[i] = key;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= M; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < (M >> 1); i += blockDim.x) {
                uint32_t pos = 2 * i - (i & (stride - 1));
                RankKey a = keys[pos], b = keys[pos + stride];
                bool up = (pos & size) == 0;
                if (key_gt(a, b) == up) {
                    keys[pos] = b;
                    keys[pos + stride] = a;
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        uint32_t m = 0;
        for (uint32_t i = 0; i < n && m < c.max_results; ++i) {
            uint32_t sk = ~(uint32_t)(keys[i].hi >> 32);
            float score = key32_to_dist(sk);
            if (score >= c.min_score) {  // retain(): NaN dropped
                order_out[m] = (uint32_t)(keys[i].lo & 0xFFFFFFFFu);
                score_out[m] = score;
                ++m;
            }
        }
        *n_out = m;
    }
}

}  // namespace

cudaError_t launch_fuse_rank(const FuseDev &f, uint32_t n, sg_fuse_weights w, sg_fuse_config c,
                             int64_t now_s, uint32_t *d_order, float *d_score, uint32_t *d_n_out,
                             float *d_all, cudaStream_t st)
{
    uint32_t M = 2;
    while (M < n) M <<= 1;
    size_t smem = (size_t)M * sizeof(RankKey);
    uint32_t threads = M / 2 < 32 ? 32 : (M / 2 > 1024 ? 1024 : M / 2);
    cudaError_t e = cudaFuncSetAttribute(fuse_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    if (e != cudaSuccess) return e;
    fuse_rank_kernel<<<1, threads, smem, st>>>(f, n, w, c, now_s, d_order, d_score, d_n_out, d_all);
    count_launch();
    return cudaGetLastError();
}

}  // namespace sg
