/*
 * shodh_oracle.c -- CPU ORACLE for the semantic-recall hot path of shodh-memory.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It restates, in plain C, the
 * arithmetic of the reference's Rust hot path so that the CUDA path can be
 * checked bit-for-bit.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  Nothing under
 * shodh_memory_b200/ links, imports or calls it.
 *
 * Build (see oracle/build.sh):
 *   gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC ...
 * -ffp-contract=off is load-bearing: rustc/LLVM never contracts `a*b+c` into an
 * fma without explicit intrinsics, so the scalar path must round every product.
 *
 * Parity pins: the reference holds no golden top-k lists or golden embeddings
 * for this path; the oracle is pinned against every
 * known-answer unit test the reference has for these functions
 * (tests/test_oracle_pins.py, each test cites the reference test it ports).
 * The reference itself cannot be compiled here (no cargo/rustc, ~600 registry
 * crates not vendored), so there is no oracle/_ref.
 *
 * All file:line citations are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SO_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* a5: dot product                                                            */
/* ------------------------------------------------------------------------- */

/* src/vector_db/distance_inline.rs:157-173 dot_product_scalar_inline.
 * What a default `cargo build` on x86-64 Linux runs (the AVX2 arm is behind a
 * compile-time cfg, :33, and no target-cpu flag is set in the repo).
 *   sum += a[i]*b[i] + a[i+1]*b[i+1] + a[i+2]*b[i+2] + a[i+3]*b[i+3]
 * parses as sum = sum + (((p0+p1)+p2)+p3), every product rounded to f32. */
SO_EXPORT float so_dot_scalar(const float *a, const float *b, uint32_t len)
{
    uint32_t unroll_len = len & ~3u;
    float sum = 0.0f;
    uint32_t i = 0;
    while (i < unroll_len) {
        float p0 = a[i] * b[i];
        float p1 = a[i + 1] * b[i + 1];
        float p2 = a[i + 2] * b[i + 2];
        float p3 = a[i + 3] * b[i + 3];
        sum = sum + (((p0 + p1) + p2) + p3);
        i += 4;
    }
    for (uint32_t j = unroll_len; j < len; ++j) {
        sum += a[j] * b[j];
    }
    return sum;
}

/* src/vector_db/distance_inline.rs:67-111 dot_product_avx2_inline, restated
 * lane by lane: one 8-lane accumulator; per 16 elements two fused
 * multiply-adds into the SAME accumulator; remaining blocks of 8; lanes summed
 * left to right (:95-103); scalar tail `result += a[j]*b[j]` (not fused). */
__attribute__((target("fma")))
SO_EXPORT float so_dot_avx2(const float *a, const float *b, uint32_t len)
{
    uint32_t simd_len = len & ~7u;
    float lane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t i = 0;
    while (i + 16 <= simd_len) {
        for (int j = 0; j < 8; ++j) lane[j] = __builtin_fmaf(a[i + j], b[i + j], lane[j]);
        for (int j = 0; j < 8; ++j) lane[j] = __builtin_fmaf(a[i + 8 + j], b[i + 8 + j], lane[j]);
        i += 16;
    }
    while (i < simd_len) {
        for (int j = 0; j < 8; ++j) lane[j] = __builtin_fmaf(a[i + j], b[i + j], lane[j]);
        i += 8;
    }
    float result = lane[0] + lane[1];
    result = result + lane[2];
    result = result + lane[3];
    result = result + lane[4];
    result = result + lane[5];
    result = result + lane[6];
    result = result + lane[7];
    for (uint32_t j = simd_len; j < len; ++j) {
        float p = a[j] * b[j];
        result = result + p;
    }
    return result;
}

/* variant: 0 = scalar build (default), 1 = avx2+fma build */
SO_EXPORT float so_dot(const float *a, const float *b, uint32_t len, int variant)
{
    return variant == 1 ? so_dot_avx2(a, b, len) : so_dot_scalar(a, b, len);
}

/* src/vector_db/distance_inline.rs:284-305 euclidean_squared_scalar_inline */
SO_EXPORT float so_euclidean_squared(const float *a, const float *b, uint32_t len)
{
    uint32_t unroll_len = len & ~3u;
    float sum = 0.0f;
    uint32_t i = 0;
    while (i < unroll_len) {
        float d0 = a[i] - b[i];
        float d1 = a[i + 1] - b[i + 1];
        float d2 = a[i + 2] - b[i + 2];
        float d3 = a[i + 3] - b[i + 3];
        float q0 = d0 * d0, q1 = d1 * d1, q2 = d2 * d2, q3 = d3 * d3;
        sum = sum + (((q0 + q1) + q2) + q3);
        i += 4;
    }
    for (uint32_t j = unroll_len; j < len; ++j) {
        float diff = a[j] - b[j];
        sum += diff * diff;
    }
    return sum;
}

/* src/vector_db/distance_inline.rs:218-252 euclidean_squared_avx2_inline, lane by lane:
 * diff = a - b; sum = fma(diff, diff, sum) per block of 8; lanes summed left to right;
 * scalar tail `result += diff*diff` (not fused). */
__attribute__((target("fma")))
SO_EXPORT float so_euclidean_squared_avx2(const float *a, const float *b, uint32_t len)
{
    uint32_t simd_len = len & ~7u;
    float lane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t i = 0; i < simd_len; i += 8)
        for (int j = 0; j < 8; ++j) {
            float diff = a[i + j] - b[i + j];
            lane[j] = __builtin_fmaf(diff, diff, lane[j]);
        }
    float result = lane[0] + lane[1];
    for (int j = 2; j < 8; ++j) result = result + lane[j];
    for (uint32_t j = simd_len; j < len; ++j) {
        float diff = a[j] - b[j];
        float p = diff * diff;
        result = result + p;
    }
    return result;
}

/* src/vector_db/distance_inline.rs:415-431 l2_norm_squared_scalar_inline */
SO_EXPORT float so_l2_norm_squared(const float *a, uint32_t len)
{
    uint32_t unroll_len = len & ~3u;
    float sum = 0.0f;
    uint32_t i = 0;
    while (i < unroll_len) {
        float q0 = a[i] * a[i], q1 = a[i + 1] * a[i + 1];
        float q2 = a[i + 2] * a[i + 2], q3 = a[i + 3] * a[i + 3];
        sum = sum + (((q0 + q1) + q2) + q3);
        i += 4;
    }
    for (uint32_t j = unroll_len; j < len; ++j) sum += a[j] * a[j];
    return sum;
}

/* src/vector_db/distance_inline.rs:357-388 l2_norm_squared_avx2_inline: sum = fma(va, va, sum) per block
 * of 8, lanes summed left to right, scalar tail not fused. */
__attribute__((target("fma")))
SO_EXPORT float so_l2_norm_squared_avx2(const float *a, uint32_t len)
{
    uint32_t simd_len = len & ~7u;
    float lane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t i = 0; i < simd_len; i += 8)
        for (int j = 0; j < 8; ++j) lane[j] = __builtin_fmaf(a[i + j], a[i + j], lane[j]);
    float result = lane[0] + lane[1];
    for (int j = 2; j < 8; ++j) result = result + lane[j];
    for (uint32_t j = simd_len; j < len; ++j) {
        float p = a[j] * a[j];
        result = result + p;
    }
    return result;
}

/* src/vector_db/distance_inline.rs:315-317 l2_norm_inline */
SO_EXPORT float so_l2_norm(const float *a, uint32_t len)
{
    return sqrtf(so_l2_norm_squared(a, len));
}

/* ------------------------------------------------------------------------- */
/* a6 / a7: distances and cosine                                              */
/* ------------------------------------------------------------------------- */

/* src/vector_db/distance_inline.rs:479-481 normalized_distance_inline = -dot */
SO_EXPORT float so_normalized_distance(const float *a, const float *b, uint32_t len, int variant)
{
    return -so_dot(a, b, len, variant);
}

/* src/similarity.rs:10-24 cosine_similarity: 0 on length mismatch or zero
 * norm; three dots; clamp to [-1,1]. */
SO_EXPORT float so_cosine_similarity(const float *a, uint32_t len_a, const float *b, uint32_t len_b,
                                     int variant)
{
    if (len_a != len_b) return 0.0f;
    float dot = so_dot(a, b, len_a, variant);
    float norm_a = sqrtf(so_dot(a, a, len_a, variant));
    float norm_b = sqrtf(so_dot(b, b, len_a, variant));
    if (norm_a == 0.0f || norm_b == 0.0f) return 0.0f;
    float c = dot / (norm_a * norm_b);
    /* f32::clamp: NaN stays NaN */
    if (c < -1.0f) c = -1.0f;
    if (c > 1.0f) c = 1.0f;
    return c;
}

/* src/vector_db/distance_inline.rs:444-458 cosine_similarity_inline:
 * dot / sqrt(|a|^2 * |b|^2), `< 1e-10 -> 0`, no clamp.  variant picks the build
 * (scalar / avx2+fma) for the dot and for both l2_norm_squared_inline calls. */
SO_EXPORT float so_cosine_similarity_inline_v(const float *a, const float *b, uint32_t len, int variant)
{
    float dot = so_dot(a, b, len, variant);
    float na = variant == 1 ? so_l2_norm_squared_avx2(a, len) : so_l2_norm_squared(a, len);
    float nb = variant == 1 ? so_l2_norm_squared_avx2(b, len) : so_l2_norm_squared(b, len);
    float denominator = sqrtf(na * nb);
    if (denominator < 1e-10f) return 0.0f;
    return dot / denominator;
}
SO_EXPORT float so_cosine_similarity_inline(const float *a, const float *b, uint32_t len)
{
    return so_cosine_similarity_inline_v(a, b, len, 0);
}

/* src/vector_db/distance_inline.rs:494-503 normalize_inplace */
SO_EXPORT void so_normalize_inplace(float *a, uint32_t len)
{
    float norm = so_l2_norm(a, len);
    if (norm > 1e-10f) {
        float inv = 1.0f / norm;
        for (uint32_t i = 0; i < len; ++i) a[i] *= inv;
    }
}

/* src/vector_db/vamana.rs:755-761 VamanaIndex::distance.
 * metric: 0 NormalizedDotProduct, 1 Euclidean, 2 Cosine */
SO_EXPORT float so_distance(const float *a, const float *b, uint32_t len, int metric, int variant)
{
    switch (metric) {
    case 1: return variant == 1 ? so_euclidean_squared_avx2(a, b, len) : so_euclidean_squared(a, b, len);
    case 2: return 1.0f - so_cosine_similarity_inline_v(a, b, len, variant);
    default: return so_normalized_distance(a, b, len, variant);
    }
}

/* ------------------------------------------------------------------------- */
/* ordering helpers                                                           */
/* ------------------------------------------------------------------------- */

/* f32::total_cmp as an order-preserving map to u32 (IEEE totalOrder:
 * -NaN < -inf < ... < -0 < +0 < ... < +inf < +NaN). */
SO_EXPORT uint32_t so_total_order_key(float x)
{
    uint32_t b;
    memcpy(&b, &x, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

typedef struct {
    uint32_t id;
    float dist;
} so_pair;

/* src/vector_db/vamana.rs:1184 / :1664-1673
 * sort_by(|a,b| a.1.total_cmp(&b.1).then_with(|| a.0.cmp(&b.0))) */
static int cmp_dist_id(const void *pa, const void *pb)
{
    const so_pair *a = (const so_pair *)pa, *b = (const so_pair *)pb;
    uint32_t ka = so_total_order_key(a->dist), kb = so_total_order_key(b->dist);
    if (ka != kb) return ka < kb ? -1 : 1;
    if (a->id != b->id) return a->id < b->id ? -1 : 1;
    return 0;
}

static inline int bitmap_get(const uint32_t *bm, uint64_t i)
{
    return bm ? (int)((bm[i >> 5] >> (i & 31)) & 1u) : 0;
}

/* ------------------------------------------------------------------------- */
/* a9: VamanaIndex::brute_force_search                                        */
/* ------------------------------------------------------------------------- */

/* src/vector_db/vamana.rs:1167-1188.  For id in 0..n, skipping deleted ids:
 * dist = distance(query,row); sort all by (dist total_cmp asc, id asc);
 * truncate k.  `as_written` != 0 also reproduces the reference's incidental
 * work (a heap clone of every row, :1179 -> :448/:496, and the full O(n log n)
 * sort) so that the "reference as written" CPU baseline is honest; with
 * as_written == 0 the same result is produced with a bounded selection and no
 * clone ("de-wasted" baseline).
 * Returns the number of results written (min(k, live rows)). */
SO_EXPORT uint32_t so_brute_force_search(const float *corpus, uint64_t n, uint32_t dim,
                                         const uint32_t *deleted_bitmap, const float *query,
                                         uint32_t k, int metric, int variant, int as_written,
                                         uint32_t *ids_out, float *dist_out)
{
    if (n == 0 || k == 0) return 0;
    if (as_written) {
        so_pair *d = (so_pair *)malloc(sizeof(so_pair) * (size_t)n);
        float *clone = (float *)malloc(sizeof(float) * dim);
        size_t m = 0;
        for (uint64_t i = 0; i < n; ++i) {
            if (bitmap_get(deleted_bitmap, i)) continue;
            memcpy(clone, corpus + i * dim, sizeof(float) * dim); /* get_vector() clone */
            d[m].id = (uint32_t)i;
            d[m].dist = so_distance(query, clone, dim, metric, variant);
            ++m;
        }
        qsort(d, m, sizeof(so_pair), cmp_dist_id);
        uint32_t r = (uint32_t)(m < k ? m : k);
        for (uint32_t j = 0; j < r; ++j) {
            ids_out[j] = d[j].id;
            dist_out[j] = d[j].dist;
        }
        free(clone);
        free(d);
        return r;
    }
    /* bounded selection: keep the k best in a sorted array (same final order) */
    so_pair *best = (so_pair *)malloc(sizeof(so_pair) * (size_t)(k + 1));
    uint32_t m = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (bitmap_get(deleted_bitmap, i)) continue;
        so_pair c;
        c.id = (uint32_t)i;
        c.dist = so_distance(query, corpus + i * dim, dim, metric, variant);
        if (m == k && cmp_dist_id(&c, &best[k - 1]) >= 0) continue;
        uint32_t pos = m < k ? m : k - 1;
        while (pos > 0 && cmp_dist_id(&c, &best[pos - 1]) < 0) {
            best[pos] = best[pos - 1];
            --pos;
        }
        best[pos] = c;
        if (m < k) ++m;
    }
    for (uint32_t j = 0; j < m; ++j) {
        ids_out[j] = best[j].id;
        dist_out[j] = best[j].dist;
    }
    free(best);
    return m;
}

/* The same search for a batch of queries, one query per host thread (the
 * reference serves concurrent recalls on tokio's blocking pool, one sequential
 * scan each: src/handlers/recall.rs:513).  ids_out/dist_out are [b,k];
 * counts_out [b].  threads <= 0 -> all. Returns threads actually used. */
SO_EXPORT int so_brute_force_search_batch(const float *corpus, uint64_t n, uint32_t dim,
                                          const uint32_t *deleted_bitmap, const float *queries,
                                          uint32_t b, uint32_t k, int metric, int variant,
                                          int as_written, int threads, uint32_t *ids_out,
                                          float *dist_out, uint32_t *counts_out)
{
    int used = 1;
#ifdef _OPENMP
    /* an explicit thread count is honoured up to the processor count: torchrun exports OMP_NUM_THREADS=1, which
     * would otherwise silently turn the all-cores baseline into a single-thread one */
    int maxt = threads > 0 ? omp_get_num_procs() : omp_get_max_threads();
    if (threads <= 0 || threads > maxt) threads = maxt;
    if ((uint32_t)threads > b) threads = (int)b;
    if (threads < 1) threads = 1;
    used = threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (uint32_t q = 0; q < b; ++q) {
        counts_out[q] = so_brute_force_search(corpus, n, dim, deleted_bitmap,
                                              queries + (size_t)q * dim, k, metric, variant,
                                              as_written, ids_out + (size_t)q * k,
                                              dist_out + (size_t)q * k);
    }
    return used;
}

/* Exact scores for a list of rows (used to spot-check full-size GPU runs). */
SO_EXPORT void so_score_rows(const float *corpus, uint32_t dim, const float *query,
                             const uint32_t *ids, uint32_t n_ids, int metric, int variant,
                             float *dist_out)
{
    for (uint32_t j = 0; j < n_ids; ++j)
        dist_out[j] = so_distance(query, corpus + (size_t)ids[j] * dim, dim, metric, variant);
}

/* ------------------------------------------------------------------------- */
/* a8: similarity.rs::top_k_similar                                           */
/* ------------------------------------------------------------------------- */

typedef struct {
    float score;
    uint32_t pos;
} so_scored;

/* OrderedFloat<f32> ordering (ordered-float crate): NaN is greater than every
 * number and equal to itself; -0.0 == +0.0. */
static int ordered_float_cmp(float a, float b)
{
    int an = isnan(a), bn = isnan(b);
    if (an || bn) return an == bn ? 0 : (an ? 1 : -1);
    return a < b ? -1 : (a > b ? 1 : 0);
}

/* src/similarity.rs:27-48: score all with cosine_similarity, stable sort by
 * score descending (ties keep input order), take k.  Candidates are rows of a
 * dense [n,dim] matrix; the generic payload T is the row position. */
SO_EXPORT uint32_t so_top_k_similar(const float *query, uint32_t qdim, const float *cands,
                                    uint64_t n, uint32_t dim, uint32_t k, int variant,
                                    uint32_t *pos_out, float *score_out)
{
    so_scored *s = (so_scored *)malloc(sizeof(so_scored) * (size_t)(n ? n : 1));
    for (uint64_t i = 0; i < n; ++i) {
        s[i].score = so_cosine_similarity(query, qdim, cands + i * dim, dim, variant);
        s[i].pos = (uint32_t)i;
    }
    /* stable insertion-style merge sort is overkill here; n is test-sized.
     * Use a stable binary-insertion sort on (score desc). */
    for (uint64_t i = 1; i < n; ++i) {
        so_scored x = s[i];
        uint64_t j = i;
        while (j > 0 && ordered_float_cmp(s[j - 1].score, x.score) < 0) {
            s[j] = s[j - 1];
            --j;
        }
        s[j] = x;
    }
    uint32_t r = (uint32_t)(n < k ? n : k);
    for (uint32_t j = 0; j < r; ++j) {
        pos_out[j] = s[j].pos;
        score_out[j] = s[j].score;
    }
    free(s);
    return r;
}

/* ------------------------------------------------------------------------- */
/* a11: RetrievalEngine::search_ids post-processing                           */
/* ------------------------------------------------------------------------- */

typedef struct {
    uint64_t mem_hi, mem_lo; /* MemoryId = Uuid, compared as 16 big-endian bytes */
    float sim;
} so_memhit;

static int cmp_memhit(const void *pa, const void *pb)
{
    const so_memhit *a = (const so_memhit *)pa, *b = (const so_memhit *)pb;
    uint32_t ka = so_total_order_key(a->sim), kb = so_total_order_key(b->sim);
    if (ka != kb) return ka > kb ? -1 : 1; /* sim desc (b.1.total_cmp(&a.1)) */
    if (a->mem_hi != b->mem_hi) return a->mem_hi < b->mem_hi ? -1 : 1;
    if (a->mem_lo != b->mem_lo) return a->mem_lo < b->mem_lo ? -1 : 1;
    return 0;
}

/* src/memory/retrieval.rs:927-961.  similarity = -distance; vector id ->
 * MemoryId (vec2mem is [n_vectors][2] u64 = uuid hi/lo; a pair of all-ones
 * means "no mapping", get_memory_id -> None); keep the max similarity per
 * memory (`if similarity > *score`); sort (similarity desc by total_cmp,
 * MemoryId asc); truncate limit. */
SO_EXPORT uint32_t so_search_ids_post(const uint32_t *vec_ids, const float *dists, uint32_t n,
                                      const uint64_t *vec2mem, uint64_t n_vectors, uint32_t limit,
                                      uint64_t *mem_out /*[limit][2]*/, float *sim_out)
{
    so_memhit *best = (so_memhit *)malloc(sizeof(so_memhit) * (size_t)(n ? n : 1));
    uint32_t m = 0;
    for (uint32_t i = 0; i < n; ++i) {
        float similarity = -dists[i];
        if (vec_ids[i] >= n_vectors) continue;
        uint64_t hi = vec2mem[2 * (uint64_t)vec_ids[i]], lo = vec2mem[2 * (uint64_t)vec_ids[i] + 1];
        if (hi == UINT64_MAX && lo == UINT64_MAX) continue;
        uint32_t j = 0;
        for (; j < m; ++j)
            if (best[j].mem_hi == hi && best[j].mem_lo == lo) break;
        if (j == m) {
            best[m].mem_hi = hi;
            best[m].mem_lo = lo;
            best[m].sim = similarity;
            ++m;
        } else if (similarity > best[j].sim) {
            best[j].sim = similarity;
        }
    }
    qsort(best, m, sizeof(so_memhit), cmp_memhit);
    uint32_t r = m < limit ? m : limit;
    for (uint32_t j = 0; j < r; ++j) {
        mem_out[2 * j] = best[j].mem_hi;
        mem_out[2 * j + 1] = best[j].mem_lo;
        sim_out[j] = best[j].sim;
    }
    free(best);
    return r;
}

/* src/memory/hybrid_search.rs:537-594 RRFusion::new + RRFusion::fuse.
 * new: weights normalised by their f32 left-fold sum (1/len each when sum <= 0).
 * fuse: lists in order, entries in rank order: scores[id] += weight / (k + (rank+1) as f32)
 * starting from 0.0 (lists beyond the weight vector use weight 1.0); the union is sorted by
 * (score total_cmp desc, MemoryId asc).  ids: [n_total][2] u64 (uuid hi/lo, big-endian halves);
 * offsets: [n_lists+1].  Returns the number of distinct ids. */
typedef struct { uint64_t hi, lo; float score; } so_rrf_ent;
static int cmp_rrf(const void *pa, const void *pb)
{
    const so_rrf_ent *a = (const so_rrf_ent *)pa, *b = (const so_rrf_ent *)pb;
    int32_t ka, kb;
    memcpy(&ka, &a->score, 4);
    memcpy(&kb, &b->score, 4);
    ka ^= (int32_t)(((uint32_t)(ka >> 31)) >> 1);     /* f32::total_cmp key */
    kb ^= (int32_t)(((uint32_t)(kb >> 31)) >> 1);
    if (ka != kb) return ka > kb ? -1 : 1;            /* descending */
    if (a->hi != b->hi) return a->hi < b->hi ? -1 : 1;
    if (a->lo != b->lo) return a->lo < b->lo ? -1 : 1;
    return 0;
}
SO_EXPORT uint32_t so_rrf_fuse(float k, const float *weights, uint32_t n_weights, const uint64_t *ids,
                               const uint32_t *offsets, uint32_t n_lists, uint64_t *ids_out, float *score_out)
{
    float wn[64];
    float sum = 0.0f;
    if (n_weights > 64) n_weights = 64;
    for (uint32_t i = 0; i < n_weights; ++i) sum = sum + weights[i];
    for (uint32_t i = 0; i < n_weights; ++i) wn[i] = sum > 0.0f ? weights[i] / sum : 1.0f / (float)n_weights;
    const uint32_t n_total = offsets[n_lists];
    so_rrf_ent *u = (so_rrf_ent *)malloc(sizeof(so_rrf_ent) * (size_t)(n_total ? n_total : 1));
    uint32_t m = 0;
    for (uint32_t l = 0; l < n_lists; ++l) {
        const float w = l < n_weights ? wn[l] : 1.0f;
        for (uint32_t e = offsets[l]; e < offsets[l + 1]; ++e) {
            const uint32_t rank = e - offsets[l];
            const float c = w / (k + (float)(rank + 1));
            uint32_t j = 0;
            for (; j < m; ++j)
                if (u[j].hi == ids[2 * (size_t)e] && u[j].lo == ids[2 * (size_t)e + 1]) break;
            if (j == m) {
                u[m].hi = ids[2 * (size_t)e];
                u[m].lo = ids[2 * (size_t)e + 1];
                u[m].score = 0.0f;
                ++m;
            }
            u[j].score = u[j].score + c;
        }
    }
    qsort(u, m, sizeof(so_rrf_ent), cmp_rrf);
    for (uint32_t j = 0; j < m; ++j) {
        ids_out[2 * (size_t)j] = u[j].hi;
        ids_out[2 * (size_t)j + 1] = u[j].lo;
        score_out[j] = u[j].score;
    }
    free(u);
    return m;
}

/* ------------------------------------------------------------------------- */
/* a12 / a13: relevance fusion                                                */
/* ------------------------------------------------------------------------- */

/* src/relevance.rs:601-606 calibrate_score; SIGMOID_STEEPNESS 10 (:97),
 * SIGMOID_MIDPOINT 0.5 (:100). */
SO_EXPORT float so_calibrate_score(float score)
{
    if (!isfinite(score)) return 0.0f;
    return 1.0f / (1.0f + expf(-10.0f * (score - 0.5f)));
}

/* src/relevance.rs:529-594 LearnedWeights::fuse_scores_full.
 * w = {semantic, entity, tag, importance, momentum, access_count, graph_strength};
 * defaults 0.18,0.17,0.05,0.05,0.28,0.14,0.13 (:64-88). */
SO_EXPORT float so_fuse_scores_full(const float *w, float semantic_score, float entity_score,
                                    float tag_score, float importance_score, float momentum_ema,
                                    uint32_t access_count, float graph_strength)
{
    float cs = so_calibrate_score(semantic_score);
    float ce = so_calibrate_score(entity_score);
    float ct = so_calibrate_score(tag_score);
    float ci = so_calibrate_score(importance_score);
    float normalized_momentum = (momentum_ema + 1.0f) / 2.0f;
    float amplified;
    if (normalized_momentum > 0.65f) {
        amplified = fminf(normalized_momentum * 1.5f, 1.0f);
    } else if (normalized_momentum < 0.40f) {
        amplified = fmaxf(normalized_momentum * 0.3f, 0.0f);
    } else {
        amplified = normalized_momentum;
    }
    float cm = so_calibrate_score(amplified);
    float access_score;
    if (access_count == 0) {
        access_score = 0.0f;
    } else {
        float log_access = log2f((float)access_count + 1.0f);
        access_score = fminf(log_access / 4.0f, 1.0f);
    }
    float ca = so_calibrate_score(access_score);
    float cg = so_calibrate_score(graph_strength);
    float result = w[0] * cs + w[1] * ce;
    result = result + w[2] * ct;
    result = result + w[3] * ci;
    result = result + w[4] * cm;
    result = result + w[5] * ca;
    result = result + w[6] * cg;
    return isfinite(result) ? result : 0.0f;
}

/* src/relevance.rs:1524-1547 apply_recency_boost.  age_hours =
 * (now - created_at).num_hours() as u64: chrono truncates toward zero and a
 * negative i64 cast to u64 becomes huge (so future timestamps get no boost). */
SO_EXPORT float so_apply_recency_boost(float base_score, int64_t now_unix_s, int64_t created_unix_s,
                                       uint64_t boost_hours, float multiplier)
{
    if (boost_hours == 0) return base_score;
    int64_t age_s = now_unix_s - created_unix_s;
    int64_t hours = age_s / 3600; /* C division truncates toward zero, like num_hours() */
    uint64_t age_hours = (uint64_t)hours;
    if (age_hours <= boost_hours) {
        float decay = 1.0f - ((float)age_hours / (float)boost_hours);
        float boost = 1.0f + (multiplier - 1.0f) * decay;
        return fminf(base_score * boost, 1.0f);
    }
    return base_score;
}

typedef struct {
    float score;
    int64_t created;
    uint32_t id;
} so_ranked;

static int cmp_ranked(const void *pa, const void *pb)
{
    const so_ranked *a = (const so_ranked *)pa, *b = (const so_ranked *)pb;
    uint32_t ka = so_total_order_key(a->score), kb = so_total_order_key(b->score);
    if (ka != kb) return ka > kb ? -1 : 1;            /* score desc, total_cmp */
    if (a->created != b->created) return a->created > b->created ? -1 : 1; /* created_at desc */
    if (a->id != b->id) return a->id < b->id ? -1 : 1; /* id asc */
    return 0;
}

/* src/relevance.rs:805-918 ranking phase of surface_relevant_inner for n
 * candidates given as SoA feature arrays: fuse (:847-855) -> recency boost
 * (:869-874; defaults 24 h, x1.2: :171-177) -> x graph_boost (1.15, :147-149)
 * capped at 1 when entity_score > 0 (:877-881) -> sort (score desc,
 * created_at desc, id asc) (:905-910) -> keep >= min_score (0.25, :914-915)
 * -> truncate max_results (:918).
 * The reference's final tie-break is on the id *string*; candidates here are
 * identified by their position, which the caller must supply in id order.
 * Returns the count; order_out holds candidate positions, score_out scores. */
SO_EXPORT uint32_t so_fuse_rank(const float *w, const float *semantic, const float *entity,
                                const float *tag, const float *importance, const float *momentum,
                                const uint32_t *access_count, const float *graph_strength,
                                const int64_t *created_unix_s, uint32_t n, int64_t now_unix_s,
                                uint64_t recency_boost_hours, float recency_multiplier,
                                float graph_boost_multiplier, float min_score, uint32_t max_results,
                                uint32_t *order_out, float *score_out, float *all_scores_out)
{
    so_ranked *r = (so_ranked *)malloc(sizeof(so_ranked) * (size_t)(n ? n : 1));
    for (uint32_t i = 0; i < n; ++i) {
        float fused = so_fuse_scores_full(w, semantic[i], entity[i], tag[i], importance[i],
                                          momentum[i], access_count[i], graph_strength[i]);
        float boosted = so_apply_recency_boost(fused, now_unix_s, created_unix_s[i],
                                               recency_boost_hours, recency_multiplier);
        float final_score = entity[i] > 0.0f ? fminf(boosted * graph_boost_multiplier, 1.0f) : boosted;
        r[i].score = final_score;
        r[i].created = created_unix_s[i];
        r[i].id = i;
        if (all_scores_out) all_scores_out[i] = final_score;
    }
    qsort(r, n, sizeof(so_ranked), cmp_ranked);
    uint32_t m = 0;
    for (uint32_t i = 0; i < n && m < max_results; ++i) {
        if (r[i].score >= min_score) { /* retain(|r| r.relevance_score >= MIN) -- NaN dropped */
            order_out[m] = r[i].id;
            score_out[m] = r[i].score;
            ++m;
        }
    }
    free(r);
    return m;
}

/* ------------------------------------------------------------------------- */
/* a3 / a4: pooling                                                           */
/* ------------------------------------------------------------------------- */

/* src/embeddings/minilm.rs:959-978 (single) / :1083-1108 (batch row): masked
 * mean-pool of last_hidden_state [seq,hidden] with attention mask [seq] (i64),
 * then finalize_pooled (:846-878, MiniLM arm: no pre-norm, no truncation):
 * NaN/Inf -> 0; norm = sqrt(sum x*x) left-to-right; divide if norm > EPSILON. */
SO_EXPORT void so_pool_finalize(const float *hidden_states, const int64_t *attention, uint32_t seq,
                                uint32_t hidden, float *out)
{
    for (uint32_t d = 0; d < hidden; ++d) out[d] = 0.0f;
    float mask_sum = 0.0f;
    for (uint32_t s = 0; s < seq; ++s) {
        if (attention[s] == 1) {
            for (uint32_t d = 0; d < hidden; ++d) out[d] += hidden_states[(size_t)s * hidden + d];
            mask_sum += 1.0f;
        }
    }
    if (mask_sum > 0.0f)
        for (uint32_t d = 0; d < hidden; ++d) out[d] /= mask_sum;
    for (uint32_t d = 0; d < hidden; ++d)
        if (isnan(out[d]) || isinf(out[d])) out[d] = 0.0f;
    float ss = 0.0f;
    for (uint32_t d = 0; d < hidden; ++d) ss += out[d] * out[d];
    float norm = sqrtf(ss);
    if (norm > 1.1920929e-7f && !isnan(norm))
        for (uint32_t d = 0; d < hidden; ++d) out[d] /= norm;
}

/* ------------------------------------------------------------------------- */
/* synthetic data -- integer-only so that the CUDA      */
/* generator (csrc/synth.cu) is bit-identical                                 */
/* ------------------------------------------------------------------------- */

static inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* approx-normal sample for (seed, stream, index): sum of four 16-bit uniforms,
 * centred, scaled by 2^-16 -- every step exact in f32. */
static inline float synth_normal(uint64_t seed, uint64_t stream, uint64_t index)
{
    uint64_t h = mix64(mix64(seed ^ (stream * 0xD1B54A32D192ED03ull)) + index);
    int32_t s = (int32_t)(h & 0xFFFF) + (int32_t)((h >> 16) & 0xFFFF) +
                (int32_t)((h >> 32) & 0xFFFF) + (int32_t)((h >> 48) & 0xFFFF) - 131070;
    return (float)s * (1.0f / 65536.0f);
}

/* row i of the corpus: x ~ approx N(0,I), L2-normalised in f32 with a
 * sequential left-to-right sum of squares and a true division. */
static void synth_row(uint64_t seed, uint64_t row, uint32_t dim, float *out)
{
    float ss = 0.0f;
    for (uint32_t j = 0; j < dim; ++j) {
        float x = synth_normal(seed, 1, row * (uint64_t)dim + j);
        out[j] = x;
        float sq = x * x;
        ss = ss + sq;
    }
    float norm = sqrtf(ss);
    if (norm > 0.0f)
        for (uint32_t j = 0; j < dim; ++j) out[j] = out[j] / norm;
}

static inline float round_through_f16(float x)
{
    _Float16 h = (_Float16)x; /* round-to-nearest-even, like __float2half_rn */
    return (float)h;
}

/* rows [first_row, first_row+n) as f32 [n,dim]; f16_round != 0 rounds every
 * element through binary16 (what an F16 store holds, upcast again). */
SO_EXPORT void so_synth_corpus(uint64_t seed, uint64_t first_row, uint64_t n, uint32_t dim,
                               int f16_round, float *out)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(omp_get_num_procs())
#endif
    for (uint64_t i = 0; i < n; ++i) {
        float *r = out + i * dim;
        synth_row(seed, first_row + i, dim, r);
        if (f16_round)
            for (uint32_t j = 0; j < dim; ++j) r[j] = round_through_f16(r[j]);
    }
}

/* queries [b,dim]: even q = fresh unit vector (stream 2); odd q =
 * normalize(row_j + 0.1*noise) for j = hash(q) mod n_rows (stream 3 noise),
 * which plants a well-separated nearest neighbour.  near_row_out[q] = j or
 * UINT64_MAX. */
SO_EXPORT void so_synth_queries(uint64_t seed, uint64_t n_rows, uint32_t b, uint32_t dim,
                                float *out, uint64_t *near_row_out)
{
    for (uint32_t q = 0; q < b; ++q) {
        float *r = out + (size_t)q * dim;
        float ss = 0.0f;
        if ((q & 1u) == 0 || n_rows == 0) {
            for (uint32_t j = 0; j < dim; ++j) r[j] = synth_normal(seed, 2, (uint64_t)q * dim + j);
            if (near_row_out) near_row_out[q] = UINT64_MAX;
        } else {
            uint64_t j0 = mix64(seed ^ (0xA5A5A5A5ull + q)) % n_rows;
            synth_row(seed, j0, dim, r);
            for (uint32_t j = 0; j < dim; ++j) {
                float noise = synth_normal(seed, 3, (uint64_t)q * dim + j);
                /* synth_normal has sigma 0.577; 0.1/sqrt(384) per element gives a
                 * perturbation of norm ~0.058 on a unit row (cos ~0.998 to row j0) */
                r[j] = r[j] + (0.1f * 0.05103104f) * noise;
            }
            if (near_row_out) near_row_out[q] = j0;
        }
        for (uint32_t j = 0; j < dim; ++j) {
            float sq = r[j] * r[j];
            ss = ss + sq;
        }
        float norm = sqrtf(ss);
        if (norm > 0.0f)
            for (uint32_t j = 0; j < dim; ++j) r[j] = r[j] / norm;
    }
}

SO_EXPORT void so_round_f16(const float *in, float *out, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i) out[i] = round_through_f16(in[i]);
}

SO_EXPORT int so_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
