// sg_dist.cuh -- the reference's distance arithmetic, operation by operation (device side).
// Shared by the exact sweep (scan_exact.cu) and the graph index kernels (vamana.cu): both must
// produce the same f32 bit patterns as src/vector_db/distance_inline.rs for a (query, row) pair.
#pragma once
#include "sg_common.cuh"

namespace sg {

// ---- per-metric accumulation of 4 consecutive elements (query q, row x) -------------------------
// METRIC 0 (NormalizedDotProduct), 2 (Cosine) and 3 (similarity.rs cosine) accumulate q*x in the
// order of dot_product_{scalar,avx2}_inline (distance_inline.rs:157-173 / :67-111); METRIC 1
// (Euclidean) accumulates (q-x)^2 in the order of euclidean_squared_{scalar,avx2}_inline
// (:284-305 / :218-252).  `o` selects the AVX lane group (element e -> lane e % 8).
template <int ORDER, int METRIC>
SG_DEVICE_INLINE void accum4(float *acc, const float4 q, float x0, float x1, float x2, float x3, int o)
{
    if constexpr (METRIC == SG_METRIC_EUCLIDEAN) {
        const float d0 = __fsub_rn(q.x, x0), d1 = __fsub_rn(q.y, x1), d2 = __fsub_rn(q.z, x2), d3 = __fsub_rn(q.w, x3);
        if constexpr (ORDER == SG_ORDER_SCALAR) {
            const float g = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)),
                                      __fmul_rn(d3, d3));
            acc[0] = __fadd_rn(acc[0], g);
        } else {
            acc[o + 0] = __fmaf_rn(d0, d0, acc[o + 0]);
            acc[o + 1] = __fmaf_rn(d1, d1, acc[o + 1]);
            acc[o + 2] = __fmaf_rn(d2, d2, acc[o + 2]);
            acc[o + 3] = __fmaf_rn(d3, d3, acc[o + 3]);
        }
    } else {
        if constexpr (ORDER == SG_ORDER_SCALAR) {
            // sum = sum + (((p0+p1)+p2)+p3), every product rounded (no FMA)
            const float g = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q.x, x0), __fmul_rn(q.y, x1)), __fmul_rn(q.z, x2)),
                                      __fmul_rn(q.w, x3));
            acc[0] = __fadd_rn(acc[0], g);
        } else {
            // 8 AVX lanes, fused multiply-add
            acc[o + 0] = __fmaf_rn(q.x, x0, acc[o + 0]);
            acc[o + 1] = __fmaf_rn(q.y, x1, acc[o + 1]);
            acc[o + 2] = __fmaf_rn(q.z, x2, acc[o + 2]);
            acc[o + 3] = __fmaf_rn(q.w, x3, acc[o + 3]);
        }
    }
}
// lanes summed left to right (distance_inline.rs:95-103); the scalar build has one accumulator
template <int ORDER>
SG_DEVICE_INLINE float hsum(const float *acc)
{
    if constexpr (ORDER == SG_ORDER_SCALAR) {
        return acc[0];
    } else {
        float r = __fadd_rn(acc[0], acc[1]);
#pragma unroll
        for (int t = 2; t < 8; ++t) r = __fadd_rn(r, acc[t]);
        return r;
    }
}
SG_DEVICE_INLINE float canon_nan(float x) { return (x != x) ? __uint_as_float(0x7FC00000u) : x; }

// accumulated sums -> the distance the ranking uses.  qn2 / rn2 = |query|^2 / |row|^2 in the same order.
template <int METRIC>
SG_DEVICE_INLINE float finish_distance(float s, float qn2, float rn2)
{
    if constexpr (METRIC == SG_METRIC_NEG_DOT) {
        return neg_dot_to_dist(s);
    } else if constexpr (METRIC == SG_METRIC_EUCLIDEAN) {
        return canon_nan(s);
    } else if constexpr (METRIC == SG_METRIC_COSINE) {
        // 1 - cosine_similarity_inline (distance_inline.rs:444-458): dot / sqrt(|a|^2 |b|^2), `< 1e-10 -> 0`
        const float den = __fsqrt_rn(__fmul_rn(qn2, rn2));
        const float c = (den < 1e-10f) ? 0.0f : __fdiv_rn(s, den);
        return canon_nan(__fsub_rn(1.0f, c));
    } else {
        // -cosine_similarity (similarity.rs:10-24): 0 on a zero norm, clamp to [-1,1] (NaN stays NaN).
        // OrderedFloat order for top_k_similar (:27-48): NaN greatest, -0 == +0, ties keep input order.
        const float na = __fsqrt_rn(qn2), nb = __fsqrt_rn(rn2);
        float c = 0.0f;
        if (!(na == 0.0f || nb == 0.0f)) {
            c = __fdiv_rn(s, __fmul_rn(na, nb));
            if (c < -1.0f) c = -1.0f;
            if (c > 1.0f) c = 1.0f;
        }
        c = (c != c) ? __uint_as_float(0x7FC00000u) : __fadd_rn(c, 0.0f);   // one NaN, no -0
        return -c;                                                            // -NaN sorts first in totalOrder
    }
}


// One (query, row) distance computed by ONE thread from a row it reads 16 bytes at a time, in the
// reference's order (the graph kernels score <= 32 scattered rows per step: one row per lane, the query
// broadcast from shared memory).  row_ptr: the row in global memory (LDG = true, read-only path) or staged
// in shared memory (LDG = false).  q_s: the query as f32 in shared memory.
template <bool F16, int ORDER, int METRIC, bool LDG>
SG_DEVICE_INLINE float row_distance_ptr(const void *row_ptr, uint32_t dim, const float *q_s, float qn2)
{
    constexpr int NACC = (ORDER == SG_ORDER_AVX2) ? 8 : 1;
    constexpr bool ROWNORM = (METRIC == SG_METRIC_COSINE || METRIC == SG_METRIC_COSINE_SIM);
    float acc[NACC], racc[NACC];
#pragma unroll
    for (int t = 0; t < NACC; ++t) acc[t] = racc[t] = 0.0f;
    if constexpr (F16) {
        const uint4 *p = static_cast<const uint4 *>(row_ptr);
        const uint32_t nchunk = dim >> 3;
#pragma unroll 8
        for (uint32_t c = 0; c < nchunk; ++c) {
            uint4 v;
            if constexpr (LDG) v = __ldg(p + c);
            else v = p[c];
            float x[8];
            float2 f;
            f = __half22float2(*reinterpret_cast<__half2 *>(&v.x)); x[0] = f.x; x[1] = f.y;
            f = __half22float2(*reinterpret_cast<__half2 *>(&v.y)); x[2] = f.x; x[3] = f.y;
            f = __half22float2(*reinterpret_cast<__half2 *>(&v.z)); x[4] = f.x; x[5] = f.y;
            f = __half22float2(*reinterpret_cast<__half2 *>(&v.w)); x[6] = f.x; x[7] = f.y;
            const float4 qa = *reinterpret_cast<const float4 *>(q_s + c * 8);
            const float4 qb = *reinterpret_cast<const float4 *>(q_s + c * 8 + 4);
            if constexpr (ROWNORM) {
                accum4<ORDER, SG_METRIC_NEG_DOT>(racc, make_float4(x[0], x[1], x[2], x[3]), x[0], x[1], x[2], x[3], 0);
                accum4<ORDER, SG_METRIC_NEG_DOT>(racc, make_float4(x[4], x[5], x[6], x[7]), x[4], x[5], x[6], x[7], 4);
            }
            accum4<ORDER, METRIC>(acc, qa, x[0], x[1], x[2], x[3], 0);
            accum4<ORDER, METRIC>(acc, qb, x[4], x[5], x[6], x[7], 4);
        }
    } else {
        const float4 *p = static_cast<const float4 *>(row_ptr);
        const uint32_t nchunk = dim >> 3;           // two float4 per step keeps the AVX lane group a constant
#pragma unroll 4
        for (uint32_t c = 0; c < nchunk; ++c) {
            float4 v0, v1;
            if constexpr (LDG) {
                v0 = __ldg(p + 2 * c);
                v1 = __ldg(p + 2 * c + 1);
            } else {
                v0 = p[2 * c];
                v1 = p[2 * c + 1];
            }
            const float4 qa = *reinterpret_cast<const float4 *>(q_s + c * 8);
            const float4 qb = *reinterpret_cast<const float4 *>(q_s + c * 8 + 4);
            if constexpr (ROWNORM) {
                accum4<ORDER, SG_METRIC_NEG_DOT>(racc, v0, v0.x, v0.y, v0.z, v0.w, 0);
                accum4<ORDER, SG_METRIC_NEG_DOT>(racc, v1, v1.x, v1.y, v1.z, v1.w, 4);
            }
            accum4<ORDER, METRIC>(acc, qa, v0.x, v0.y, v0.z, v0.w, 0);
            accum4<ORDER, METRIC>(acc, qb, v1.x, v1.y, v1.z, v1.w, 4);
        }
    }
    float rn2 = 0.0f;
    if constexpr (ROWNORM) rn2 = hsum<ORDER>(racc);
    return finish_distance<METRIC>(hsum<ORDER>(acc), qn2, rn2);
}

template <bool F16, int ORDER, int METRIC>
SG_DEVICE_INLINE float row_distance(const void *rows, uint32_t dim, uint64_t row, const float *q_s, float qn2)
{
    const uint8_t *p = static_cast<const uint8_t *>(rows) + row * dim * (F16 ? 2 : 4);
    return row_distance_ptr<F16, ORDER, METRIC, true>(p, dim, q_s, qn2);
}

// |v|^2 of an f32 vector in shared memory, in the order of l2_norm_squared_*_inline (== dot(v, v))
template <int ORDER>
SG_DEVICE_INLINE float norm2_smem(const float *v, uint32_t dim)
{
    constexpr int NACC = (ORDER == SG_ORDER_AVX2) ? 8 : 1;
    float nacc[NACC];
#pragma unroll
    for (int t = 0; t < NACC; ++t) nacc[t] = 0.0f;
    for (uint32_t d = 0; d < dim; d += 8) {
        const float4 a = *reinterpret_cast<const float4 *>(v + d);
        const float4 b = *reinterpret_cast<const float4 *>(v + d + 4);
        accum4<ORDER, SG_METRIC_NEG_DOT>(nacc, a, a.x, a.y, a.z, a.w, 0);
        accum4<ORDER, SG_METRIC_NEG_DOT>(nacc, b, b.x, b.y, b.z, b.w, 4);
    }
    return hsum<ORDER>(nacc);
}

}  // namespace sg
