// graph_api.cu -- C ABI of the Vamana graph index over the GPU store (kernels: vamana.cu).
//
// Mirrors the graph half of the reference's VamanaIndex (src/vector_db/vamana.rs): build :200-284,
// search's default ANN arm :779-807 (greedy_search :576-657), add_vector :853-973, plus upload /
// download of a graph so an index built by the reference (or loaded from its VAMA file) is searched
// as-is.  The exact sweep (sg_store_search) needs none of this; a graph exists only after
// sg_store_build_graph / sg_store_set_graph / sg_store_enable_graph / a VAMA load.
#include <cfloat>
#include <cstdlib>
#include <cstring>

#include "sg_store_internal.h"

namespace {

constexpr uint32_t GRAPH_NONE = 0xFFFFFFFFu;
constexpr uint32_t MAX_DEGREE_LIMIT = 64;
constexpr uint32_t MAX_BEAM = 2048;

uint32_t pow2_at_least(uint64_t x)
{
    uint32_t p = 1;
    while (p < x && p < (1u << 30)) p <<= 1;
    return p;
}

int ensure_graph_rows(sg_store *s, uint64_t need)
{
    if (need <= s->graph_cap) return SG_OK;
    uint64_t ncap = s->graph_cap + s->graph_cap / 2;
    if (ncap < need) ncap = need;
    if (ncap < s->cap) ncap = s->cap;        // follow the row matrix so appends rarely reallocate
    uint32_t *ng = nullptr;
    SG_CUDA_TRY(cudaMalloc(&ng, ncap * (size_t)s->gs * 4));
    cudaError_t e = cudaMemsetAsync(ng, 0xFF, ncap * (size_t)s->gs * 4, s->stream);
    if (e == cudaSuccess && s->d_graph && s->graph_len)
        e = cudaMemcpyAsync(ng, s->d_graph, s->graph_len * (size_t)s->gs * 4, cudaMemcpyDeviceToDevice, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) {
        cudaFree(ng);
        set_error("growing the graph failed: %s", cudaGetErrorString(e));
        return SG_ERR_CUDA;
    }
    if (s->d_graph) cudaFree(s->d_graph);
    s->d_graph = ng;
    s->graph_cap = ncap;
    return SG_OK;
}

int check_graph_config(const sg_store *s, uint32_t max_degree, uint32_t search_list_size, float alpha)
{
    if (s->metric == SG_METRIC_COSINE_SIM) {
        set_error("the graph index takes the VamanaIndex metrics (NormalizedDotProduct, Euclidean, Cosine)");
        return SG_ERR_UNSUPPORTED;
    }
    if (max_degree == 0 || max_degree > MAX_DEGREE_LIMIT) {
        set_error("max_degree=%u outside 1..%u", max_degree, MAX_DEGREE_LIMIT);
        return SG_ERR_UNSUPPORTED;
    }
    if (search_list_size == 0 || search_list_size + max_degree > 128) {
        set_error("search_list_size=%u: search_list_size + max_degree must be <= 128", search_list_size);
        return SG_ERR_UNSUPPORTED;
    }
    if (!(alpha >= 1.0f) || alpha > 4.0f) {
        set_error("alpha=%f outside 1..4", (double)alpha);
        return SG_ERR_INVALID;
    }
    return SG_OK;
}

uint32_t default_table_size(uint32_t K)
{
    uint32_t h = pow2_at_least((uint64_t)K * 64);
    if (h < 4096) h = 4096;
    if (h > (1u << 20)) h = 1u << 20;
    return h;
}

// b beam searches on st.  Exactly one of d_q / d_query_rows is set.  K = beam, k_out = results per query.
// d_flags / d_stats may be null.  The visited tables live in s->d_vis (grown here, reset per call).
int ann_launch(sg_store *s, const float *d_q, const uint32_t *d_query_rows, uint32_t b, uint32_t K, uint32_t k_out,
               bool use_deleted, uint32_t id_base, uint64_t n_rows, uint64_t graph_len, uint32_t H, uint32_t *d_ids,
               float *d_dist, uint32_t *d_counts, uint32_t *d_flags, unsigned long long *d_stats, cudaStream_t st)
{
    // bound the visited tables of one launch to 256 MB: larger batches go in pieces
    uint32_t piece = (uint32_t)(((size_t)64 << 20) / H);
    if (piece == 0) piece = 1;
    if (piece > b) piece = b;
    int rc = grow(s->d_vis, s->vis_cap, (size_t)piece * H);
    if (rc != SG_OK) return rc;
    for (uint32_t done = 0; done < b; done += piece) {
        const uint32_t nq = b - done < piece ? b - done : piece;
        SG_CUDA_TRY(cudaMemsetAsync(s->d_vis, 0xFF, (size_t)nq * H * 4, st));
        AnnArgs a;
        a.rows = s->d_rows;
        a.dim = s->dim;
        a.n_rows = n_rows;
        a.graph = s->d_graph;
        a.gs = s->gs;
        a.graph_len = graph_len;
        a.entry = s->medoid;
        a.queries = d_q ? d_q + (size_t)done * s->dim : nullptr;
        a.query_rows = d_query_rows ? d_query_rows + done : nullptr;
        a.nq = nq;
        a.K = K;
        a.k_out = k_out;
        a.deleted = (use_deleted && s->n_deleted) ? s->d_deleted : nullptr;
        a.id_base = id_base;
        a.vis = s->d_vis;
        a.H = H;
        a.ids_out = d_ids + (size_t)done * k_out;
        a.dist_out = d_dist + (size_t)done * k_out;
        a.counts_out = d_counts + done;
        a.flags = d_flags ? d_flags + done : nullptr;
        a.stats = d_stats ? d_stats + 2 * (size_t)done : nullptr;
        a.phase = s->d_ann_phase;
        cudaError_t e = launch_ann_search(a, s->dtype == SG_DTYPE_F16, s->order, s->metric, st);
        if (e != cudaSuccess) {
            set_error("graph search launch failed: %s", cudaGetErrorString(e));
            return SG_ERR_CUDA;
        }
    }
    return SG_OK;
}

// VamanaIndex::search preamble (:765-796): empty index -> no results; no graph -> error; widen the beam by
// min(deleted, 2k) when anything is soft-deleted
int ann_beam(const sg_store *s, uint32_t k, uint32_t *K_out)
{
    if (k == 0 || k > 1024) {
        set_error("k=%u outside the supported range 1..1024", k);
        return k == 0 ? SG_ERR_INVALID : SG_ERR_UNSUPPORTED;
    }
    if (s->n && (s->gs == 0 || s->graph_len == 0)) {
        set_error("Vamana graph not built. Call sg_store_build_graph / sg_store_set_graph first.");   // vamana.rs:780-784
        return SG_ERR_NOT_BUILT;
    }
    uint64_t K = k;
    if (s->n_deleted) K += s->n_deleted < 2ull * k ? s->n_deleted : 2ull * k;
    if (K > MAX_BEAM) K = MAX_BEAM;
    *K_out = (uint32_t)K;
    return SG_OK;
}

}  // namespace

namespace sg {

void graph_free(sg_store *s)
{
    cudaFree(s->d_graph);
    cudaFree(s->d_vis);
    cudaFree(s->d_ann_flags);
    cudaFree(s->d_ann_stats);
    cudaFree(s->d_g32);
    cudaFree(s->d_gf32);
    cudaFree(s->d_ann_phase);
    s->d_graph = nullptr;
    s->d_vis = nullptr;
}

// VamanaIndex::add_vector (vamana.rs:853-973) for rows [first, first+n), one after the other, all on the
// device and stream-ordered: beam search with k = max_degree over the graph as it stands, then the link
// kernel (new list + back-edges, closest max_degree kept).  The graph equals the oracle's add_vector
// sequence edge for edge.
int graph_append_locked(sg_store *s, uint64_t first, uint64_t n)
{
    if (s->gs == 0 || n == 0) return SG_OK;
    int rc = ensure_graph_rows(s, first + n);
    if (rc != SG_OK) return rc;
    const uint32_t gs = s->gs;
    // scratch: [0] the node id, [1] result count, [2..2+gs) result ids ; f32: result distances
    if ((rc = grow(s->d_g32, s->g32_cap, (size_t)2 + gs)) != SG_OK) return rc;
    if ((rc = grow(s->d_gf32, s->gf32_cap, (size_t)gs)) != SG_OK) return rc;
    const uint32_t H = default_table_size(gs);
    cudaStream_t st = s->stream;
    for (uint64_t id = first; id < first + n; ++id) {
        if (id == 0 || s->graph_len == 0) {
            // the first vector: a node with no neighbours, and it is the medoid (:896-906)
            SG_CUDA_TRY(cudaMemsetAsync(s->d_graph + id * gs, 0xFF, (size_t)gs * 4, st));
            if (id == 0) s->medoid = 0;
            s->graph_len = id + 1;
            continue;
        }
        SG_CUDA_TRY(launch_iota(s->d_g32, (uint32_t)id, 1, st));
        rc = ann_launch(s, nullptr, s->d_g32, 1, gs, gs, false, 0, id + 1, s->graph_len, H, s->d_g32 + 2, s->d_gf32,
                        s->d_g32 + 1, nullptr, nullptr, st);
        if (rc != SG_OK) return rc;
        cudaError_t e = launch_graph_link(s->d_rows, s->dtype == SG_DTYPE_F16, s->order, s->metric, s->dim, s->d_graph, gs,
                                          (uint32_t)id, s->d_g32 + 2, s->d_g32 + 1, st);
        if (e != cudaSuccess) {
            set_error("graph link launch failed: %s", cudaGetErrorString(e));
            return SG_ERR_CUDA;
        }
        s->graph_len = id + 1;
    }
    SG_CUDA_TRY(cudaStreamSynchronize(st));
    return SG_OK;
}

}  // namespace sg

extern "C" {

int sg_store_enable_graph(sg_store *s, uint32_t max_degree, uint32_t search_list_size, float alpha)
{
    if (!s) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    int rc = check_graph_config(s, max_degree, search_list_size, alpha);
    if (rc != SG_OK) return rc;
    if (s->gs != 0 || s->n != 0) {
        set_error("sg_store_enable_graph needs an empty store without a graph (use sg_store_build_graph for existing rows)");
        return SG_ERR_INVALID;
    }
    s->gs = max_degree;
    s->search_list_size = search_list_size;
    s->alpha = alpha;
    s->graph_len = 0;
    s->medoid = 0;
    return SG_OK;
}

int sg_store_graph_info(sg_store *s, uint32_t *max_degree_out, uint32_t *medoid_out, uint64_t *graph_len_out)
{
    if (!s) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    if (max_degree_out) *max_degree_out = s->gs;
    if (medoid_out) *medoid_out = s->medoid;
    if (graph_len_out) *graph_len_out = s->gs ? s->graph_len : 0;
    return SG_OK;
}

int sg_store_set_graph(sg_store *s, uint32_t max_degree, uint32_t medoid, const uint32_t *deg, const uint32_t *nbr,
                       uint32_t stride, uint64_t n_nodes)
{
    if (!s || (n_nodes && (!deg || !nbr))) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    int rc = check_graph_config(s, max_degree, 1, 1.0f);
    if (rc != SG_OK) return rc;
    if (n_nodes != s->n) {
        set_error("the graph has %llu nodes, the store %llu rows", (unsigned long long)n_nodes, (unsigned long long)s->n);
        return SG_ERR_INVALID;
    }
    if (n_nodes && medoid >= n_nodes) {
        set_error("medoid %u out of range", medoid);
        return SG_ERR_INVALID;
    }
    for (uint64_t i = 0; i < n_nodes; ++i)
        if (deg[i] > max_degree || deg[i] > stride) {
            set_error("node %llu lists %u neighbours (max_degree %u, stride %u)", (unsigned long long)i, deg[i], max_degree, stride);
            return SG_ERR_INVALID;
        }
    if (s->gs != max_degree) {             // a different list stride: start over
        cudaFree(s->d_graph);
        s->d_graph = nullptr;
        s->graph_cap = 0;
        s->graph_len = 0;
        s->gs = max_degree;
    }
    s->graph_len = 0;
    if ((rc = ensure_graph_rows(s, n_nodes ? n_nodes : 1)) != SG_OK) return rc;
    std::vector<uint32_t> packed;
    const uint64_t piece = 1u << 20;       // nodes per upload
    for (uint64_t i0 = 0; i0 < n_nodes; i0 += piece) {
        const uint64_t m = n_nodes - i0 < piece ? n_nodes - i0 : piece;
        packed.assign((size_t)m * max_degree, GRAPH_NONE);
        for (uint64_t i = 0; i < m; ++i)
            memcpy(packed.data() + i * max_degree, nbr + (i0 + i) * stride, (size_t)deg[i0 + i] * 4);
        SG_CUDA_TRY(cudaMemcpy(s->d_graph + i0 * max_degree, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice));
    }
    s->medoid = medoid;
    s->graph_len = n_nodes;
    return SG_OK;
}

int sg_store_get_graph(sg_store *s, uint32_t *deg_out, uint32_t *nbr_out, uint64_t n_nodes)
{
    if (!s || (n_nodes && (!deg_out || !nbr_out))) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    if (n_nodes == 0 && (s->gs == 0 || s->graph_len == 0)) return SG_OK;
    if (s->gs == 0 || n_nodes != s->graph_len) {
        set_error("sg_store_get_graph: the graph has %llu nodes, %llu requested", (unsigned long long)(s->gs ? s->graph_len : 0),
                  (unsigned long long)n_nodes);
        return SG_ERR_INVALID;
    }
    if (n_nodes == 0) return SG_OK;
    int rc = grow(s->d_g32, s->g32_cap, (size_t)n_nodes);
    if (rc != SG_OK) return rc;
    SG_CUDA_TRY(launch_graph_degree(s->d_graph, s->gs, n_nodes, s->d_g32, s->stream));
    SG_CUDA_TRY(cudaMemcpyAsync(deg_out, s->d_g32, (size_t)n_nodes * 4, cudaMemcpyDeviceToHost, s->stream));
    SG_CUDA_TRY(cudaMemcpyAsync(nbr_out, s->d_graph, (size_t)n_nodes * s->gs * 4, cudaMemcpyDeviceToHost, s->stream));
    SG_CUDA_TRY(cudaStreamSynchronize(s->stream));
    return SG_OK;
}

int sg_store_build_graph(sg_store *s, uint32_t max_degree, uint32_t search_list_size, float alpha, uint64_t seed)
{
    if (!s) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    int rc = check_graph_config(s, max_degree, search_list_size, alpha);
    if (rc != SG_OK) return rc;
    if (s->gs != max_degree) {
        cudaFree(s->d_graph);
        s->d_graph = nullptr;
        s->graph_cap = 0;
    }
    s->gs = max_degree;
    s->search_list_size = search_list_size;
    s->alpha = alpha;
    s->graph_len = 0;
    s->medoid = 0;
    const uint64_t n = s->n;
    if (n == 0) return SG_OK;                                  // build(vec![]) is a no-op (:201-203)
    if ((rc = ensure_graph_rows(s, n)) != SG_OK) return rc;
    cudaStream_t st = s->stream;
    const uint32_t gs = max_degree, L = search_list_size;
    const bool f16 = s->dtype == SG_DTYPE_F16;

    // Step 1 (:212, :287-316): random start graph
    SG_CUDA_TRY(launch_graph_init(s->d_graph, gs, n, seed, st));

    // Step 3 (:218, :400-441): medoid = the row closest to the centroid -- centroid summed in row order by
    // centroid_kernel, the arg-min by the exact sweep (ties -> smallest id, as the reference's `<` scan)
    uint64_t B = n / 32;
    if (B < 64) B = 64;
    if (B > 16384) B = 16384;
    if (const char *e = getenv("SG_GRAPH_BATCH")) B = (uint64_t)atoll(e) > 0 ? (uint64_t)atoll(e) : B;   // tuning knobs
    if (B > n) B = n;
    int passes = 2;
    if (const char *e = getenv("SG_GRAPH_PASSES")) passes = atoi(e) > 0 ? atoi(e) : 2;
    const size_t o_nodes = 0, o_cnt = o_nodes + B, o_res = o_cnt + B, o_new = o_res + B * L, o_touch = o_new + B * gs,
                 o_ntouch = o_touch + B * gs, g32_total = o_ntouch + 4;
    if ((rc = grow(s->d_g32, s->g32_cap, g32_total)) != SG_OK) return rc;
    if ((rc = grow(s->d_gf32, s->gf32_cap, B * L + s->dim + 8)) != SG_OK) return rc;
    float *d_centroid = s->d_gf32 + B * L;
    SG_CUDA_TRY(launch_centroid(s->d_rows, f16, s->dim, n, d_centroid, st));
    {
        uint32_t *d_mid = s->d_g32, *d_mcnt = s->d_g32 + 1;
        float *d_mdist = s->d_gf32;
        const uint32_t saved_base = s->id_base;
        s->id_base = 0;
        rc = search_exact_locked(s, d_centroid, 1, 1, d_mid, d_mdist, d_mcnt, st, false, false);
        s->id_base = saved_base;
        if (rc != SG_OK) return rc;
        uint32_t h_id = 0, h_cnt = 0;
        float h_dist = 0.0f;
        SG_CUDA_TRY(cudaMemcpyAsync(&h_id, d_mid, 4, cudaMemcpyDeviceToHost, st));
        SG_CUDA_TRY(cudaMemcpyAsync(&h_cnt, d_mcnt, 4, cudaMemcpyDeviceToHost, st));
        SG_CUDA_TRY(cudaMemcpyAsync(&h_dist, d_mdist, 4, cudaMemcpyDeviceToHost, st));
        SG_CUDA_TRY(cudaStreamSynchronize(st));
        s->medoid = (h_cnt && h_dist < FLT_MAX) ? h_id : 0;      // `dist < best_dist`, best_dist = f32::MAX (:428-437)
    }
    s->graph_len = n;

    // Step 4 (:221-281): two passes; per batch: beam search (L) from the medoid -> robust_prune -> edges + back-edges
    uint32_t ecap = 32;
    if (const char *e = getenv("SG_GRAPH_INBOX")) ecap = (uint32_t)atoi(e);
    if (ecap < 1) ecap = 1;
    if (ecap + gs > 128) ecap = 128 - gs;
    uint32_t *d_inbox = nullptr, *d_inbox_cnt = nullptr;
    SG_CUDA_TRY(cudaMalloc(&d_inbox, n * (size_t)ecap * 4));
    cudaError_t e = cudaMalloc(&d_inbox_cnt, n * 4);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_inbox_cnt, 0, n * 4, st);
    if (e != cudaSuccess) {
        cudaFree(d_inbox);
        cudaFree(d_inbox_cnt);
        set_error("graph build scratch: %s", cudaGetErrorString(e));
        return SG_ERR_OOM;
    }
    const uint32_t H = default_table_size(L);
    // Back-edge prunes order their candidates by true distance.  The reference hands that prune zero distances, so it
    // walks the list in id order (:251-262); one insertion at a time that is harmless, but with a whole batch of
    // arrivals per target it systematically drops the highest ids: recall@10 0.87 vs 1.00 on 300k x 384 (beam 40).
    // SG_GRAPH_BACKEDGE_BY_ID=1 selects the reference's order.
    const char *bd = getenv("SG_GRAPH_BACKEDGE_BY_ID");
    const bool by_id = bd && *bd == '1';
    uint32_t *d_nodes = s->d_g32 + o_nodes, *d_cnt = s->d_g32 + o_cnt, *d_res = s->d_g32 + o_res, *d_new = s->d_g32 + o_new,
             *d_touch = s->d_g32 + o_touch, *d_ntouch = s->d_g32 + o_ntouch;
    rc = SG_OK;
    for (int pass = 0; pass < passes && rc == SG_OK; ++pass) {
        for (uint64_t first = 0; first < n && rc == SG_OK; first += B) {
            const uint32_t m = (uint32_t)(n - first < B ? n - first : B);
            e = launch_iota(d_nodes, (uint32_t)first, m, st);
            if (e == cudaSuccess)
                rc = ann_launch(s, nullptr, d_nodes, m, L, L, false, 0, n, n, H, d_res, s->d_gf32, d_cnt, nullptr, nullptr, st);
            if (rc != SG_OK) break;
            if (e == cudaSuccess)
                e = launch_build_prune(s->d_rows, f16, s->dim, s->metric, alpha, s->d_graph, gs, d_nodes, m, d_res, d_cnt, L,
                                       d_new, st);
            if (e == cudaSuccess) e = cudaMemsetAsync(d_ntouch, 0, 4, st);
            if (e == cudaSuccess)
                e = launch_build_apply(s->d_graph, gs, d_nodes, m, d_new, d_inbox, d_inbox_cnt, ecap, d_touch, d_ntouch, st);
            if (e == cudaSuccess) {
                uint64_t grid = (uint64_t)m * gs;
                const uint64_t gmax = (uint64_t)s->sm_count * 8;
                if (grid > gmax) grid = gmax;
                e = launch_build_backedges(s->d_rows, f16, s->dim, s->metric, alpha, s->d_graph, gs, d_touch, d_ntouch, d_inbox,
                                           d_inbox_cnt, ecap, by_id, (uint32_t)grid, st);
            }
            if (e != cudaSuccess) {
                set_error("graph build launch failed: %s", cudaGetErrorString(e));
                rc = SG_ERR_CUDA;
            }
        }
    }
    e = cudaStreamSynchronize(st);
    cudaFree(d_inbox);
    cudaFree(d_inbox_cnt);
    if (rc == SG_OK && e != cudaSuccess) {
        set_error("graph build failed: %s", cudaGetErrorString(e));
        rc = SG_ERR_CUDA;
    }
    if (rc != SG_OK) s->graph_len = 0;
    return rc;
}

int sg_store_search_ann_device(sg_store *s, const float *d_q, uint32_t b, uint32_t k, uint32_t *d_ids, float *d_dist,
                               uint32_t *d_counts, uint32_t *d_flags, void *cuda_stream)
{
    if (!s || (b && (!d_q || !d_ids || !d_dist || !d_counts))) return SG_ERR_INVALID;
    if (b == 0) return SG_OK;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    uint32_t K = 0;
    int rc = ann_beam(s, k, &K);
    if (rc != SG_OK) return rc;
    if ((rc = chain_begin(s, st)) != SG_OK) return rc;         // the visited tables are per-store scratch
    if (s->n == 0) {                                           // Ok(vec![]) (:766-768)
        SG_CUDA_TRY(launch_merge_keys(nullptr, b, 0, k, s->id_base, nullptr, d_ids, d_dist, d_counts, st));
        if (d_flags) SG_CUDA_TRY(cudaMemsetAsync(d_flags, 0, (size_t)b * 4, st));
        return chain_end(s, st);
    }
    s->ann_searches += b;
    rc = ann_launch(s, d_q, nullptr, b, K, k, true, s->id_base, s->n, s->graph_len, default_table_size(K), d_ids, d_dist,
                    d_counts, d_flags, nullptr, st);
    const int rc2 = chain_end(s, st);
    return rc != SG_OK ? rc : rc2;
}

int sg_store_search_ann(sg_store *s, const float *queries, uint32_t b, uint32_t k, uint32_t *ids_out, float *dist_out,
                        uint32_t *counts_out, uint64_t *stats_out)
{
    if (!s || (b && (!queries || !ids_out || !dist_out || !counts_out))) return SG_ERR_INVALID;
    if (b == 0) return SG_OK;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    uint32_t K = 0;
    int rc = ann_beam(s, k, &K);
    if (rc != SG_OK) return rc;
    if (s->n == 0) {
        for (uint32_t q = 0; q < b; ++q) counts_out[q] = 0;
        for (size_t i = 0; i < (size_t)b * k; ++i) {
            ids_out[i] = GRAPH_NONE;
            dist_out[i] = __builtin_inff();
        }
        if (stats_out) memset(stats_out, 0, (size_t)b * 16);
        return SG_OK;
    }
    const size_t qbytes = (size_t)b * s->dim * 4, rbytes = (size_t)b * k * 4, cbytes = (size_t)b * 4, sbytes = (size_t)b * 16;
    if ((rc = chain_begin(s, s->stream)) != SG_OK) return rc;
    if ((rc = grow_pinned(s, qbytes + 2 * rbytes + 2 * cbytes + sbytes)) != SG_OK) return rc;
    if ((rc = grow(s->d_queries, s->q_cap, (size_t)b * s->dim)) != SG_OK) return rc;
    if ((size_t)b * k > s->res_cap) {
        size_t c1 = s->res_cap, c2 = s->res_cap;
        if ((rc = grow(s->d_ids, c1, (size_t)b * k)) != SG_OK) return rc;
        if ((rc = grow(s->d_dist, c2, (size_t)b * k)) != SG_OK) return rc;
        s->res_cap = c1 < c2 ? c1 : c2;
    }
    if ((rc = grow(s->d_counts, s->cnt_cap, (size_t)b)) != SG_OK) return rc;
    if ((rc = grow(s->d_ann_flags, s->ann_flags_cap, (size_t)b)) != SG_OK) return rc;
    if ((rc = grow(s->d_ann_stats, s->ann_stats_cap, (size_t)b * 2)) != SG_OK) return rc;
    cudaStream_t st = s->stream;
    uint8_t *hp = static_cast<uint8_t *>(s->h_pinned);
    uint8_t *h_ids = hp + qbytes, *h_dist = h_ids + rbytes, *h_cnt = h_dist + rbytes, *h_flags = h_cnt + cbytes,
            *h_stats = h_flags + cbytes;
    memcpy(hp, queries, qbytes);
    SG_CUDA_TRY(cudaMemcpyAsync(s->d_queries, hp, qbytes, cudaMemcpyHostToDevice, st));
    const bool prof = s->profiling && 2 * (size_t)(s->ev_count + 1) <= s->ev.size();
    if (prof) SG_CUDA_TRY(cudaEventRecord(s->ev[2 * s->ev_count], st));
    uint32_t H = default_table_size(K);
    rc = ann_launch(s, s->d_queries, nullptr, b, K, k, true, s->id_base, s->n, s->graph_len, H, s->d_ids, s->d_dist, s->d_counts,
                    s->d_ann_flags, reinterpret_cast<unsigned long long *>(s->d_ann_stats), st);
    if (rc != SG_OK) return rc;
    if (prof) {
        SG_CUDA_TRY(cudaEventRecord(s->ev[2 * s->ev_count + 1], st));
        ++s->ev_count;
    }
    SG_CUDA_TRY(cudaMemcpyAsync(h_ids, s->d_ids, rbytes, cudaMemcpyDeviceToHost, st));
    SG_CUDA_TRY(cudaMemcpyAsync(h_dist, s->d_dist, rbytes, cudaMemcpyDeviceToHost, st));
    SG_CUDA_TRY(cudaMemcpyAsync(h_cnt, s->d_counts, cbytes, cudaMemcpyDeviceToHost, st));
    SG_CUDA_TRY(cudaMemcpyAsync(h_flags, s->d_ann_flags, cbytes, cudaMemcpyDeviceToHost, st));
    SG_CUDA_TRY(cudaMemcpyAsync(h_stats, s->d_ann_stats, sbytes, cudaMemcpyDeviceToHost, st));
    SG_CUDA_TRY(cudaStreamSynchronize(st));
    s->ann_searches += b;
    // a query whose visited table filled up is run again, alone, with a table 8x the size (rare: the default
    // holds 32 visited ids per beam slot)
    const uint32_t *flags = reinterpret_cast<const uint32_t *>(h_flags);
    for (uint32_t q = 0; q < b; ++q) {
        uint32_t f = flags[q];
        uint32_t Hq = H;
        while ((f & 1u) && Hq < (1u << 26)) {
            Hq <<= 3;
            ++s->ann_retries;
            rc = ann_launch(s, s->d_queries + (size_t)q * s->dim, nullptr, 1, K, k, true, s->id_base, s->n, s->graph_len, Hq,
                            s->d_ids + (size_t)q * k, s->d_dist + (size_t)q * k, s->d_counts + q, s->d_ann_flags + q,
                            reinterpret_cast<unsigned long long *>(s->d_ann_stats) + 2 * (size_t)q, st);
            if (rc != SG_OK) return rc;
            SG_CUDA_TRY(cudaMemcpyAsync(h_ids + (size_t)q * k * 4, s->d_ids + (size_t)q * k, (size_t)k * 4, cudaMemcpyDeviceToHost, st));
            SG_CUDA_TRY(cudaMemcpyAsync(h_dist + (size_t)q * k * 4, s->d_dist + (size_t)q * k, (size_t)k * 4, cudaMemcpyDeviceToHost, st));
            SG_CUDA_TRY(cudaMemcpyAsync(h_cnt + (size_t)q * 4, s->d_counts + q, 4, cudaMemcpyDeviceToHost, st));
            SG_CUDA_TRY(cudaMemcpyAsync(h_flags + (size_t)q * 4, s->d_ann_flags + q, 4, cudaMemcpyDeviceToHost, st));
            SG_CUDA_TRY(cudaMemcpyAsync(h_stats + (size_t)q * 16, s->d_ann_stats + 2 * (size_t)q, 16, cudaMemcpyDeviceToHost, st));
            SG_CUDA_TRY(cudaStreamSynchronize(st));
            f = flags[q];
        }
    }
    memcpy(ids_out, h_ids, rbytes);
    memcpy(dist_out, h_dist, rbytes);
    memcpy(counts_out, h_cnt, cbytes);
    if (stats_out) memcpy(stats_out, h_stats, sbytes);
    return SG_OK;
}

int sg_store_ann_phase_cycles(sg_store *s, int enable, uint64_t *cycles_out)
{
    if (!s) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    if (s->d_ann_phase && cycles_out) {
        SG_CUDA_TRY(cudaStreamSynchronize(s->stream));
        SG_CUDA_TRY(cudaMemcpy(cycles_out, s->d_ann_phase, 8 * 8, cudaMemcpyDeviceToHost));
    } else if (cycles_out) {
        memset(cycles_out, 0, 8 * 8);
    }
    if (enable && !s->d_ann_phase) SG_CUDA_TRY(cudaMalloc(&s->d_ann_phase, 8 * 8));
    if (!enable && s->d_ann_phase) {
        cudaFree(s->d_ann_phase);
        s->d_ann_phase = nullptr;
    }
    if (s->d_ann_phase) SG_CUDA_TRY(cudaMemset(s->d_ann_phase, 0, 8 * 8));
    return SG_OK;
}

int sg_store_ann_stats(sg_store *s, uint64_t *queries_out, uint64_t *retries_out)
{
    if (!s) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    if (queries_out) *queries_out = s->ann_searches;
    if (retries_out) *retries_out = s->ann_retries;
    return SG_OK;
}

}  // extern "C"
