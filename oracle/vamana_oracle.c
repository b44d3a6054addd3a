/*
 * vamana_oracle.c -- CPU ORACLE for the Vamana graph index.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (see shodh_oracle.c for the
 * rules: only tests/, smoke() and bench.py's CPU legs may load it).
 *
 * Plain-C restatement of reference src/vector_db/vamana.rs:
 *   greedy_search      :576-657     VamanaIndex::search (ANN arm)  :779-807
 *   robust_prune       :665-746     find_medoid                    :400-441
 *   build              :200-284     initialize_graph               :287-316
 *   add_vector         :853-973
 * SearchCandidate is ordered by (distance total_cmp, id) (:1663-1673), a TOTAL
 * order, so the two BinaryHeaps of greedy_search can be any priority queue: the
 * element popped is uniquely determined.  Float comparisons that the reference
 * makes with `>` / `<` on f32 (not total_cmp) are kept as float comparisons.
 *
 * The graph is passed as flat arrays: deg[n] and nbr[n][stride] (stride >=
 * max_degree + 1: build/add_vector let a list reach max_degree + 1 before it is
 * pruned back).  initialize_graph uses rand::thread_rng in the reference, i.e.
 * it is NOT reproducible there; the oracle draws the same kind of sample
 * (`degree` distinct non-self ids) from a seeded generator.
 *
 * Parity pins: vamana.rs tests :1685-1711 (construction + search), :1744-1774
 * (estimate_recall >= 0.6) are ported in tests/test_vamana_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SO_EXPORT __attribute__((visibility("default")))

float so_distance(const float *a, const float *b, uint32_t len, int metric, int variant);
uint32_t so_total_order_key(float x);

typedef struct {
    uint32_t id;
    float dist;
} vo_cand;

/* (distance total_cmp, id) -- vamana.rs:1663-1673 */
static int vo_cmp(const vo_cand *a, const vo_cand *b)
{
    uint32_t ka = so_total_order_key(a->dist), kb = so_total_order_key(b->dist);
    if (ka != kb) return ka < kb ? -1 : 1;
    if (a->id != b->id) return a->id < b->id ? -1 : 1;
    return 0;
}
static int vo_cmp_q(const void *a, const void *b) { return vo_cmp((const vo_cand *)a, (const vo_cand *)b); }

/* binary min-heap on vo_cmp (the `candidates` heap of Reverse<SearchCandidate>) */
typedef struct {
    vo_cand *a;
    size_t n, cap;
} vo_heap;
static void heap_push(vo_heap *h, vo_cand c)
{
    if (h->n == h->cap) {
        h->cap = h->cap ? h->cap * 2 : 256;
        h->a = (vo_cand *)realloc(h->a, h->cap * sizeof(vo_cand));
    }
    size_t i = h->n++;
    while (i > 0) {
        size_t p = (i - 1) / 2;
        if (vo_cmp(&c, &h->a[p]) >= 0) break;
        h->a[i] = h->a[p];
        i = p;
    }
    h->a[i] = c;
}
static vo_cand heap_pop(vo_heap *h)
{
    vo_cand top = h->a[0], last = h->a[--h->n];
    size_t i = 0;
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, m = i;
        const vo_cand *best = &last;
        if (l < h->n && vo_cmp(&h->a[l], best) < 0) { m = l; best = &h->a[l]; }
        if (r < h->n && vo_cmp(&h->a[r], best) < 0) { m = r; best = &h->a[r]; }
        if (m == i) break;
        h->a[i] = h->a[m];
        i = m;
    }
    if (h->n) h->a[i] = last;
    return top;
}

/* vamana.rs:576-657 greedy_search(query, k, entry).  `w` (max-heap bounded to k) is kept
 * as a sorted array.  graph_len = graph.len() (:606: nodes beyond it are skipped).
 * stats_out (may be NULL): [0] = nodes expanded, [1] = distance evaluations.
 * Returns the number of results (ascending by (dist, id)). */
SO_EXPORT uint32_t so_greedy_search(const float *corpus, uint64_t n, uint32_t dim, const uint32_t *deg,
                                    const uint32_t *nbr, uint32_t stride, uint64_t graph_len,
                                    const float *query, uint32_t k, uint32_t entry, int metric, int variant,
                                    uint32_t *ids_out, float *dist_out, uint64_t *stats_out)
{
    if (k == 0 || n == 0 || entry >= n) return 0;
    uint8_t *visited = (uint8_t *)calloc((size_t)n, 1);
    vo_heap cand = {0, 0, 0};
    vo_cand *w = (vo_cand *)malloc(sizeof(vo_cand) * ((size_t)k + 1));
    uint32_t wn = 0;
    uint64_t hops = 0, evals = 1;

    vo_cand e;
    e.id = entry;
    e.dist = so_distance(query, corpus + (size_t)entry * dim, dim, metric, variant);
    heap_push(&cand, e);
    w[wn++] = e;
    visited[entry] = 1;

    while (cand.n) {
        vo_cand cur = heap_pop(&cand);
        if (wn && cur.dist > w[wn - 1].dist) break;              /* :597-603, float `>` */
        if ((uint64_t)cur.id >= graph_len) continue;             /* :606-609 */
        ++hops;
        const uint32_t *nb = nbr + (size_t)cur.id * stride;
        for (uint32_t j = 0; j < deg[cur.id]; ++j) {
            uint32_t v = nb[j];
            if (v >= n || visited[v]) continue;                  /* (out-of-range ids would be an Err in the reference) */
            visited[v] = 1;
            vo_cand c;
            c.id = v;
            c.dist = so_distance(query, corpus + (size_t)v * dim, dim, metric, variant);
            ++evals;
            int should_add = wn < k || c.dist < w[wn - 1].dist;  /* :625, float `<` */
            if (!should_add) continue;
            heap_push(&cand, c);
            uint32_t pos = wn;                                   /* w.push + (len > k -> pop the max) */
            while (pos > 0 && vo_cmp(&c, &w[pos - 1]) < 0) {
                w[pos] = w[pos - 1];
                --pos;
            }
            w[pos] = c;
            ++wn;
            if (wn > k) --wn;
        }
    }
    for (uint32_t i = 0; i < wn; ++i) {
        ids_out[i] = w[i].id;
        dist_out[i] = w[i].dist;
    }
    if (stats_out) {
        stats_out[0] = hops;
        stats_out[1] = evals;
    }
    free(w);
    free(cand.a);
    free(visited);
    return wn;
}

static inline int vo_bit(const uint32_t *bm, uint64_t i) { return bm ? (int)((bm[i >> 5] >> (i & 31)) & 1u) : 0; }

/* vamana.rs:764-807 VamanaIndex::search, ANN arm: search_k = k + min(deleted_count, 2k) when
 * anything is deleted; greedy_search from the medoid; drop deleted ids; take k. */
SO_EXPORT uint32_t so_vamana_search(const float *corpus, uint64_t n, uint32_t dim, const uint32_t *deg,
                                    const uint32_t *nbr, uint32_t stride, uint32_t medoid,
                                    const uint32_t *deleted_bitmap, uint64_t deleted_count, const float *query,
                                    uint32_t k, int metric, int variant, uint32_t *ids_out, float *dist_out,
                                    uint64_t *stats_out)
{
    if (n == 0 || k == 0) return 0;
    uint64_t extra = deleted_count < 2ull * k ? deleted_count : 2ull * k;
    uint32_t search_k = deleted_count ? (uint32_t)(k + extra) : k;
    uint32_t *ids = (uint32_t *)malloc(sizeof(uint32_t) * search_k);
    float *dist = (float *)malloc(sizeof(float) * search_k);
    uint32_t r = so_greedy_search(corpus, n, dim, deg, nbr, stride, n, query, search_k, medoid, metric, variant,
                                  ids, dist, stats_out);
    uint32_t m = 0;
    for (uint32_t i = 0; i < r && m < k; ++i) {
        if (vo_bit(deleted_bitmap, ids[i])) continue;
        ids_out[m] = ids[i];
        dist_out[m] = dist[i];
        ++m;
    }
    free(ids);
    free(dist);
    return m;
}

/* vamana.rs:665-746 robust_prune(node_id, candidates): sort by (distance, id); drop the node
 * itself; keep a candidate unless some already-kept e has alpha*(d(c,e)+off) <= d(n,c)+off and
 * d(c,e) <= d(n,e); stop at max_degree.  off = 1 for NormalizedDotProduct, else 0 (:711-727). */
SO_EXPORT uint32_t so_robust_prune(const float *corpus, uint32_t dim, uint32_t node_id, const uint32_t *cand_ids,
                                   const float *cand_dist, uint32_t n_cand, uint32_t max_degree, float alpha,
                                   int metric, int variant, uint32_t *out_ids)
{
    if (n_cand == 0) return 0;
    vo_cand *c = (vo_cand *)malloc(sizeof(vo_cand) * n_cand);
    for (uint32_t i = 0; i < n_cand; ++i) {
        c[i].id = cand_ids[i];
        c[i].dist = cand_dist[i];
    }
    qsort(c, n_cand, sizeof(vo_cand), vo_cmp_q);
    const float *node = corpus + (size_t)node_id * dim;
    const float off = metric == 0 ? 1.0f : 0.0f;
    float *ne = (float *)malloc(sizeof(float) * (max_degree ? max_degree : 1));
    uint32_t m = 0;
    for (uint32_t i = 0; i < n_cand && m < max_degree; ++i) {
        if (c[i].id == node_id) continue;
        const float *cv = corpus + (size_t)c[i].id * dim;
        float dist_nc = so_distance(node, cv, dim, metric, variant);
        int add = 1;
        for (uint32_t e = 0; e < m; ++e) {
            float dist_ce = so_distance(cv, corpus + (size_t)out_ids[e] * dim, dim, metric, variant);
            float lhs = alpha * (dist_ce + off), rhs = dist_nc + off;
            if (lhs <= rhs && dist_ce <= ne[e]) {
                add = 0;
                break;
            }
        }
        if (add) {
            out_ids[m] = c[i].id;
            ne[m] = dist_nc;
            ++m;
        }
    }
    free(ne);
    free(c);
    return m;
}

/* vamana.rs:400-441 find_medoid: centroid summed row by row in f32, divided by n; the first
 * row with the strictly smallest distance(row, centroid) below f32::MAX wins (else 0). */
SO_EXPORT uint32_t so_find_medoid(const float *corpus, uint64_t n, uint32_t dim, int metric, int variant)
{
    if (n == 0) return 0;
    float *cen = (float *)calloc(dim, sizeof(float));
    for (uint64_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < dim; ++j) cen[j] += corpus[i * dim + j];
    for (uint32_t j = 0; j < dim; ++j) cen[j] /= (float)n;
    uint32_t best = 0;
    float best_dist = 3.402823466e+38f;
    for (uint64_t i = 0; i < n; ++i) {
        float d = so_distance(corpus + i * dim, cen, dim, metric, variant);
        if (d < best_dist) {
            best_dist = d;
            best = (uint32_t)i;
        }
    }
    free(cen);
    return best;
}

static uint64_t vo_mix(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* vamana.rs:287-316 initialize_graph: every node gets min(max_degree, n-1) distinct random
 * neighbours, never itself (sampled index idx over n-1 slots maps to idx + (idx >= i)). */
SO_EXPORT void so_vamana_init_graph(uint64_t n, uint32_t max_degree, uint32_t stride, uint64_t seed,
                                    uint32_t *deg, uint32_t *nbr)
{
    uint32_t degree = (uint32_t)(n > 0 && n - 1 < max_degree ? n - 1 : max_degree);
    if (n == 0) return;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t *row = nbr + i * stride;
        uint32_t m = 0;
        uint64_t ctr = 0;
        while (m < degree) {
            uint64_t idx = vo_mix(seed ^ vo_mix(i * 0x10001ull + (ctr++))) % (n - 1);
            uint32_t j = (uint32_t)(idx >= i ? idx + 1 : idx);
            int dup = 0;
            for (uint32_t t = 0; t < m; ++t) dup |= row[t] == j;
            if (!dup) row[m++] = j;
        }
        deg[i] = degree;
    }
}

static int row_contains(const uint32_t *row, uint32_t d, uint32_t v)
{
    for (uint32_t t = 0; t < d; ++t)
        if (row[t] == v) return 1;
    return 0;
}

/* vamana.rs:200-284 build (after initialize_graph / store_vectors / find_medoid): at most two
 * passes over the nodes in id order; per node: greedy_search(L) from the medoid, robust_prune,
 * replace the list if it changed and push back-edges, pruning a neighbour that exceeds
 * max_degree with its own list as zero-distance candidates (:251-262).  deg/nbr hold the
 * initial graph on entry and the built graph on return.  Returns the number of passes run. */
SO_EXPORT uint32_t so_vamana_build(const float *corpus, uint64_t n, uint32_t dim, uint32_t max_degree,
                                   uint32_t search_list_size, float alpha, int metric, int variant,
                                   uint32_t medoid, uint32_t stride, uint32_t *deg, uint32_t *nbr)
{
    if (n == 0 || stride < max_degree + 1) return 0;
    uint32_t L = search_list_size;
    uint32_t *ids = (uint32_t *)malloc(sizeof(uint32_t) * (L ? L : 1));
    float *dist = (float *)malloc(sizeof(float) * (L ? L : 1));
    uint32_t *pruned = (uint32_t *)malloc(sizeof(uint32_t) * (stride + 1));
    float *zero = (float *)calloc(stride + 1, sizeof(float));
    uint32_t iteration = 0;
    for (;;) {
        ++iteration;
        uint64_t updates = 0;
        for (uint64_t node = 0; node < n; ++node) {
            uint32_t r = so_greedy_search(corpus, n, dim, deg, nbr, stride, n, corpus + node * dim, L, medoid,
                                          metric, variant, ids, dist, NULL);
            uint32_t pn = so_robust_prune(corpus, dim, (uint32_t)node, ids, dist, r, max_degree, alpha, metric,
                                          variant, pruned);
            uint32_t *row = nbr + node * stride;
            int same = deg[node] == pn && memcmp(row, pruned, sizeof(uint32_t) * pn) == 0;
            if (same) continue;
            ++updates;
            memcpy(row, pruned, sizeof(uint32_t) * pn);
            deg[node] = pn;
            for (uint32_t t = 0; t < pn; ++t) {
                uint32_t v = pruned[t];
                if (v >= n) continue;
                uint32_t *vr = nbr + (size_t)v * stride;
                if (row_contains(vr, deg[v], (uint32_t)node)) continue;
                vr[deg[v]++] = (uint32_t)node;
                if (deg[v] > max_degree) {
                    uint32_t tmp[64 + 2];
                    uint32_t *out = stride + 1 <= 66 ? tmp : (uint32_t *)malloc(sizeof(uint32_t) * (stride + 1));
                    uint32_t m = so_robust_prune(corpus, dim, v, vr, zero, deg[v], max_degree, alpha, metric,
                                                 variant, out);
                    memcpy(vr, out, sizeof(uint32_t) * m);
                    deg[v] = m;
                    if (out != tmp) free(out);
                }
            }
        }
        if (updates == 0 || iteration >= 2) break;
    }
    free(zero);
    free(pruned);
    free(dist);
    free(ids);
    return iteration;
}

/* vamana.rs:853-973 add_vector for the row that was just appended as id = n_before (the corpus
 * already holds it): neighbours = greedy_search(vector, max_degree, medoid) as they come (no
 * robust_prune); each neighbour gets the back-edge and, beyond max_degree, keeps its max_degree
 * closest by (distance, id) (:938-958).  The first vector becomes the medoid with no edges. */
SO_EXPORT void so_vamana_add_vector(const float *corpus, uint64_t n_before, uint32_t dim, uint32_t max_degree,
                                    int metric, int variant, uint32_t medoid, uint32_t stride, uint32_t *deg,
                                    uint32_t *nbr)
{
    const uint32_t id = (uint32_t)n_before;
    uint32_t *row = nbr + (size_t)id * stride;
    deg[id] = 0;
    if (n_before == 0) return;
    uint32_t *ids = (uint32_t *)malloc(sizeof(uint32_t) * (max_degree ? max_degree : 1));
    float *dist = (float *)malloc(sizeof(float) * (max_degree ? max_degree : 1));
    /* the search runs over the graph as it was (graph.len() == n_before) but storage already holds the row */
    uint32_t r = so_greedy_search(corpus, n_before + 1, dim, deg, nbr, stride, n_before, corpus + (size_t)id * dim,
                                  max_degree, medoid, metric, variant, ids, dist, NULL);
    memcpy(row, ids, sizeof(uint32_t) * r);
    deg[id] = r;
    for (uint32_t t = 0; t < r; ++t) {
        uint32_t v = ids[t];
        uint32_t *vr = nbr + (size_t)v * stride;
        vr[deg[v]++] = id;
        if (deg[v] > max_degree) {
            vo_cand *c = (vo_cand *)malloc(sizeof(vo_cand) * deg[v]);
            for (uint32_t j = 0; j < deg[v]; ++j) {
                c[j].id = vr[j];
                c[j].dist = so_distance(corpus + (size_t)v * dim, corpus + (size_t)vr[j] * dim, dim, metric, variant);
            }
            qsort(c, deg[v], sizeof(vo_cand), vo_cmp_q);
            for (uint32_t j = 0; j < max_degree; ++j) vr[j] = c[j].id;
            deg[v] = max_degree;
            free(c);
        }
    }
    free(dist);
    free(ids);
}

/* vamana.rs estimate_recall-style check used by the tests: fraction of the exact top-k
 * (brute force, same metric) that the ANN search returns, averaged over the given queries. */
SO_EXPORT double so_recall_at_k(const uint32_t *ann_ids, const uint32_t *ann_counts, const uint32_t *exact_ids,
                                const uint32_t *exact_counts, uint32_t b, uint32_t k)
{
    double hit = 0, tot = 0;
    for (uint32_t q = 0; q < b; ++q) {
        for (uint32_t i = 0; i < exact_counts[q]; ++i) {
            uint32_t want = exact_ids[(size_t)q * k + i];
            for (uint32_t j = 0; j < ann_counts[q]; ++j)
                if (ann_ids[(size_t)q * k + j] == want) {
                    hit += 1;
                    break;
                }
            tot += 1;
        }
    }
    return tot > 0 ? hit / tot : 1.0;
}
