/*
 * spann_oracle.c -- CPU ORACLE for the SPANN partitioned index.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (see shodh_oracle.c for the rules).
 *
 * Plain-C restatement of reference src/vector_db/spann.rs:
 *   compute_distance      :562-572  Euclidean: sum (x-y)^2 ; otherwise 1 - sum x*y  (f32 Iterator::sum, left to right)
 *   find_nearest_centroid :545-558  first centroid with the strictly smallest distance below f32::MAX
 *   kmeans_cluster        :466-542  (the shuffled index order is an INPUT: thread_rng in the reference);
 *                                   assignments start at 0, early exit when nothing changed
 *   build                 :363-463  partitions = ceil(sqrt(n)) unless configured (:136-139); entries in vector order
 *   search                :574-693  rank partitions by (distance total_cmp, index), probe the first num_probes,
 *                                   ADC every entry (pq.rs distance_with_table), keep the k smallest in a max-heap
 *                                   of (OrderedFloat(dist), id), final sort by (distance total_cmp, id)
 *   insert                :1006-1051
 * The posting lists are passed as CSR: part_off[P+1], then ids / codes of all entries grouped by partition.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SO_EXPORT __attribute__((visibility("default")))

uint32_t so_total_order_key(float x);
void so_pq_table(const float *centroids, uint32_t m, uint32_t k, uint32_t ds, const float *query, float *table_out);
float so_pq_distance_with_table(const float *table, uint32_t m, uint32_t k, const uint8_t *codes);

/* spann.rs:562-572; metric: 0 NormalizedDotProduct, 1 Euclidean, 2 Cosine */
SO_EXPORT float so_spann_distance(const float *a, const float *b, uint32_t dim, int metric)
{
    float s = 0.0f;
    if (metric == 1) {
        for (uint32_t i = 0; i < dim; ++i) {
            float d = a[i] - b[i];
            float p = d * d;
            s = s + p;
        }
        return s;
    }
    for (uint32_t i = 0; i < dim; ++i) {
        float p = a[i] * b[i];
        s = s + p;
    }
    return 1.0f - s;
}

/* spann.rs:545-558 */
SO_EXPORT uint32_t so_spann_nearest_centroid(const float *v, const float *centroids, uint32_t k, uint32_t dim, int metric)
{
    uint32_t best = 0;
    float best_d = 3.402823466e+38f;
    for (uint32_t c = 0; c < k; ++c) {
        float d = so_spann_distance(v, centroids + (size_t)c * dim, dim, metric);
        if (d < best_d) {
            best_d = d;
            best = c;
        }
    }
    return best;
}

/* spann.rs:466-542.  Returns the number of iterations run. */
SO_EXPORT uint32_t so_spann_kmeans(const float *vectors, uint64_t n, uint32_t dim, uint32_t k, uint32_t iterations, int metric,
                                   const uint32_t *order, float *cent)
{
    for (uint32_t c = 0; c < k; ++c) {
        uint64_t src = c < n ? order[c] : order[c % n];
        memcpy(cent + (size_t)c * dim, vectors + src * dim, sizeof(float) * dim);
    }
    uint32_t *assign = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
    float *acc = (float *)malloc(sizeof(float) * (size_t)k * dim);
    uint64_t *cnt = (uint64_t *)malloc(sizeof(uint64_t) * k);
    uint32_t it = 0;
    for (; it < iterations; ++it) {
        uint64_t changed = 0;
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t a = so_spann_nearest_centroid(vectors + i * dim, cent, k, dim, metric);
            if (a != assign[i]) ++changed;
            assign[i] = a;
        }
        memset(acc, 0, sizeof(float) * (size_t)k * dim);
        memset(cnt, 0, sizeof(uint64_t) * k);
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t c = assign[i];
            ++cnt[c];
            for (uint32_t j = 0; j < dim; ++j) acc[(size_t)c * dim + j] += vectors[i * dim + j];
        }
        for (uint32_t c = 0; c < k; ++c)
            if (cnt[c] > 0)
                for (uint32_t j = 0; j < dim; ++j) cent[(size_t)c * dim + j] = acc[(size_t)c * dim + j] / (float)cnt[c];
        if (changed == 0) {
            ++it;
            break;
        }
    }
    free(cnt);
    free(acc);
    free(assign);
    return it;
}

/* build step 3 (:412-432): partition of every vector */
SO_EXPORT void so_spann_assign(const float *vectors, uint64_t n, uint32_t dim, const float *centroids, uint32_t k, int metric,
                               uint32_t *part_out)
{
    for (uint64_t i = 0; i < n; ++i) part_out[i] = so_spann_nearest_centroid(vectors + i * dim, centroids, k, dim, metric);
}

typedef struct {
    uint32_t idx;
    float dist;
} sp_rank;
static int sp_cmp(const void *pa, const void *pb)
{
    const sp_rank *a = (const sp_rank *)pa, *b = (const sp_rank *)pb;
    uint32_t ka = so_total_order_key(a->dist), kb = so_total_order_key(b->dist);
    if (ka != kb) return ka < kb ? -1 : 1;
    if (a->idx != b->idx) return a->idx < b->idx ? -1 : 1;
    return 0;
}
/* (OrderedFloat(dist), id) tuple order: NaN greatest and equal to itself, -0 == +0 */
static int sp_heap_cmp(const sp_rank *a, const sp_rank *b)
{
    int an = isnan(a->dist), bn = isnan(b->dist);
    if (an || bn) {
        if (an != bn) return an ? 1 : -1;
    } else if (a->dist != b->dist) {
        return a->dist < b->dist ? -1 : 1;
    }
    if (a->idx != b->idx) return a->idx < b->idx ? -1 : 1;
    return 0;
}

/* spann.rs:574-693 (in-memory arm).  probes_out (may be NULL) receives the probed partition ids. */
SO_EXPORT uint32_t so_spann_search(const float *centroids, uint32_t n_part, uint32_t dim, int metric, uint32_t num_probes,
                                   const uint64_t *part_off, const uint32_t *ids, const uint8_t *codes,
                                   const float *pq_centroids, uint32_t m, uint32_t kc, uint32_t ds, const float *query,
                                   uint32_t topk, uint32_t *ids_out, float *dist_out, uint32_t *probes_out)
{
    if (n_part == 0 || topk == 0) return 0;
    sp_rank *pr = (sp_rank *)malloc(sizeof(sp_rank) * n_part);
    for (uint32_t p = 0; p < n_part; ++p) {
        pr[p].idx = p;
        pr[p].dist = so_spann_distance(query, centroids + (size_t)p * dim, dim, metric);
    }
    qsort(pr, n_part, sizeof(sp_rank), sp_cmp);
    uint32_t np = num_probes < n_part ? num_probes : n_part;
    float *table = (float *)malloc(sizeof(float) * (size_t)m * kc);
    so_pq_table(pq_centroids, m, kc, ds, query, table);
    sp_rank *heap = (sp_rank *)malloc(sizeof(sp_rank) * ((size_t)topk + 1));   /* kept sorted ascending by the heap's order */
    uint32_t hn = 0;
    for (uint32_t t = 0; t < np; ++t) {
        uint32_t p = pr[t].idx;
        if (probes_out) probes_out[t] = p;
        for (uint64_t e = part_off[p]; e < part_off[p + 1]; ++e) {
            sp_rank c;
            c.idx = ids[e];
            c.dist = so_pq_distance_with_table(table, m, kc, codes + e * m);
            /* push, then pop the largest when over k */
            uint32_t pos = hn;
            while (pos > 0 && sp_heap_cmp(&c, &heap[pos - 1]) < 0) {
                heap[pos] = heap[pos - 1];
                --pos;
            }
            heap[pos] = c;
            if (hn < topk) ++hn;
        }
    }
    qsort(heap, hn, sizeof(sp_rank), sp_cmp);
    for (uint32_t i = 0; i < hn; ++i) {
        ids_out[i] = heap[i].idx;
        dist_out[i] = heap[i].dist;
    }
    free(heap);
    free(table);
    free(pr);
    return hn;
}
