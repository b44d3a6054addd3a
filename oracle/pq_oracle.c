/*
 * pq_oracle.c -- CPU ORACLE for Product Quantization.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (see shodh_oracle.c for the rules).
 *
 * Plain-C restatement of reference src/vector_db/pq.rs:
 *   squared_l2_distance :393-401   (x - y).powi(2) summed left to right (f32 Iterator::sum)
 *   kmeans              :152-215   (the shuffled index order is an INPUT: the reference draws it
 *                                   from rand::thread_rng, so it is not reproducible there)
 *   fit                 :114-149   one k-means per subspace, k = min(num_centroids, n)
 *   encode              :220-257   first centroid with the strictly smallest distance below f32::MAX
 *   decode              :260-290
 *   build_distance_table :329-353, distance_with_table :358-369 (f32::MAX for a code out of range)
 *   CompressedVectorStore::search :449-466  sort by (distance total_cmp, id), truncate k
 * Centroids are held as one flat array [num_subvectors][num_centroids][subvec_dim].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SO_EXPORT __attribute__((visibility("default")))

uint32_t so_total_order_key(float x);

/* pq.rs:393-401 */
static float pq_sq_l2(const float *a, const float *b, uint32_t n)
{
    float s = 0.0f;
    for (uint32_t i = 0; i < n; ++i) {
        float d = a[i] - b[i];
        float p = d * d; /* powi(2) */
        s = s + p;
    }
    return s;
}

/* pq.rs:152-215 kmeans over n vectors of `ds` floats with stride `stride` (the subspace slice of the
 * training matrix); order[] = the shuffled indices.  centroids_out [k][ds]. */
static void pq_kmeans(const float *vecs, uint64_t n, uint32_t stride, uint32_t ds, uint32_t k, uint32_t iterations,
                      const uint32_t *order, float *cent)
{
    for (uint32_t c = 0; c < k; ++c) {
        uint64_t src = c < n ? order[c] : order[c % n];
        memcpy(cent + (size_t)c * ds, vecs + src * stride, sizeof(float) * ds);
    }
    uint32_t *assign = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
    float *acc = (float *)malloc(sizeof(float) * (size_t)k * ds);
    uint64_t *cnt = (uint64_t *)malloc(sizeof(uint64_t) * k);
    for (uint32_t it = 0; it < iterations; ++it) {
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t best = 0;
            float best_d = 3.402823466e+38f;
            for (uint32_t c = 0; c < k; ++c) {
                float d = pq_sq_l2(vecs + i * stride, cent + (size_t)c * ds, ds);
                if (d < best_d) {
                    best_d = d;
                    best = c;
                }
            }
            assign[i] = best;
        }
        memset(acc, 0, sizeof(float) * (size_t)k * ds);
        memset(cnt, 0, sizeof(uint64_t) * k);
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t c = assign[i];
            ++cnt[c];
            for (uint32_t j = 0; j < ds; ++j) acc[(size_t)c * ds + j] += vecs[i * stride + j];
        }
        for (uint32_t c = 0; c < k; ++c)
            if (cnt[c] > 0)
                for (uint32_t j = 0; j < ds; ++j) cent[(size_t)c * ds + j] = acc[(size_t)c * ds + j] / (float)cnt[c];
    }
    free(cnt);
    free(acc);
    free(assign);
}

/* ProductQuantizer::train / fit (:94-149).  order [m][n]: one shuffled index order per subspace.
 * Returns k = min(num_centroids, n); centroids_out [m][k][ds]. */
SO_EXPORT uint32_t so_pq_train(const float *vectors, uint64_t n, uint32_t dim, uint32_t ds, uint32_t num_centroids,
                               uint32_t iterations, const uint32_t *order, float *centroids_out)
{
    if (n == 0 || ds == 0 || dim % ds) return 0;
    const uint32_t m = dim / ds;
    const uint32_t k = (uint64_t)num_centroids < n ? num_centroids : (uint32_t)n;
    for (uint32_t s = 0; s < m; ++s)
        pq_kmeans(vectors + (size_t)s * ds, n, dim, ds, k, iterations, order + (size_t)s * n,
                  centroids_out + (size_t)s * k * ds);
    return k;
}

/* encode (:220-257) for n vectors -> codes [n][m] */
SO_EXPORT void so_pq_encode(const float *centroids, uint32_t m, uint32_t k, uint32_t ds, const float *vectors,
                            uint64_t n, uint8_t *codes_out)
{
    const uint32_t dim = m * ds;
    for (uint64_t i = 0; i < n; ++i)
        for (uint32_t s = 0; s < m; ++s) {
            uint8_t best = 0;
            float best_d = 3.402823466e+38f;
            for (uint32_t c = 0; c < k; ++c) {
                float d = pq_sq_l2(vectors + i * dim + (size_t)s * ds, centroids + ((size_t)s * k + c) * ds, ds);
                if (d < best_d) {
                    best_d = d;
                    best = (uint8_t)c;
                }
            }
            codes_out[i * m + s] = best;
        }
}

/* decode (:260-290) */
SO_EXPORT void so_pq_decode(const float *centroids, uint32_t m, uint32_t k, uint32_t ds, const uint8_t *codes,
                            uint64_t n, float *vectors_out)
{
    for (uint64_t i = 0; i < n; ++i)
        for (uint32_t s = 0; s < m; ++s)
            memcpy(vectors_out + (i * m + s) * ds, centroids + ((size_t)s * k + codes[i * m + s]) * ds, sizeof(float) * ds);
}

/* build_distance_table (:329-353) -> table [m][k] */
SO_EXPORT void so_pq_table(const float *centroids, uint32_t m, uint32_t k, uint32_t ds, const float *query, float *table_out)
{
    for (uint32_t s = 0; s < m; ++s)
        for (uint32_t c = 0; c < k; ++c)
            table_out[(size_t)s * k + c] = pq_sq_l2(query + (size_t)s * ds, centroids + ((size_t)s * k + c) * ds, ds);
}

/* distance_with_table (:358-369) */
SO_EXPORT float so_pq_distance_with_table(const float *table, uint32_t m, uint32_t k, const uint8_t *codes)
{
    float total = 0.0f;
    for (uint32_t s = 0; s < m; ++s) {
        if (codes[s] >= k) return 3.402823466e+38f;
        total += table[(size_t)s * k + codes[s]];
    }
    return total;
}

/* asymmetric_distance (:296-323) */
SO_EXPORT float so_pq_asymmetric_distance(const float *centroids, uint32_t m, uint32_t k, uint32_t ds, const float *query,
                                          const uint8_t *codes)
{
    float total = 0.0f;
    for (uint32_t s = 0; s < m; ++s)
        total += pq_sq_l2(query + (size_t)s * ds, centroids + ((size_t)s * k + codes[s]) * ds, ds);
    return total;
}

typedef struct {
    uint32_t id;
    float dist;
} pq_hit;
static int pq_cmp(const void *pa, const void *pb)
{
    const pq_hit *a = (const pq_hit *)pa, *b = (const pq_hit *)pb;
    uint32_t ka = so_total_order_key(a->dist), kb = so_total_order_key(b->dist);
    if (ka != kb) return ka < kb ? -1 : 1;
    if (a->id != b->id) return a->id < b->id ? -1 : 1;
    return 0;
}

/* CompressedVectorStore::search (:449-466): table, one distance per stored vector, sort by
 * (distance total_cmp, id), truncate k.  ids [n] (unique), codes [n][m]. */
SO_EXPORT uint32_t so_pq_search(const float *centroids, uint32_t m, uint32_t k, uint32_t ds, const uint32_t *ids,
                                const uint8_t *codes, uint64_t n, const float *query, uint32_t topk, uint32_t *ids_out,
                                float *dist_out)
{
    float *table = (float *)malloc(sizeof(float) * (size_t)m * k);
    so_pq_table(centroids, m, k, ds, query, table);
    pq_hit *h = (pq_hit *)malloc(sizeof(pq_hit) * (n ? n : 1));
    for (uint64_t i = 0; i < n; ++i) {
        h[i].id = ids[i];
        h[i].dist = so_pq_distance_with_table(table, m, k, codes + i * m);
    }
    qsort(h, n, sizeof(pq_hit), pq_cmp);
    uint32_t r = (uint32_t)(n < topk ? n : topk);
    for (uint32_t i = 0; i < r; ++i) {
        ids_out[i] = h[i].id;
        dist_out[i] = h[i].dist;
    }
    free(h);
    free(table);
    return r;
}
