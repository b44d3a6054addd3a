// sg_internal.h -- declarations shared between the translation units of libshodh_gpu.
#pragma once
#include <mutex>

#include "sg_common.cuh"

namespace sg {

// "dynamic shared memory opt-in already raised to" per (kernel, device): a process-wide table that two stores on two
// host threads may reach together, hence the lock.  One DevAttr per kernel instantiation.
struct DevAttr {
    std::mutex mu;
    size_t set[64] = {0};
};
template <typename K>
inline cudaError_t ensure_smem(K kern, DevAttr &da, size_t smem)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> g(da.mu);
    size_t &cur = da.set[dev & 63];
    if (smem > cur) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cur = smem;
    }
    return cudaSuccess;
}

// ---- scan_exact.cu ----
struct ScanArgs {
    const float *queries;     // device [nq][dim] f32
    uint32_t nq;              // queries in this sweep (<= QB of the instantiation)
    uint32_t dim;
    uint64_t n_rows;
    const uint32_t *deleted;  // device bitmap (bit i = row i soft-deleted) or nullptr
    uint32_t k;
    uint32_t cap;             // per-(warp,query) candidate buffer capacity (pow2, >= k + 32)
    uint32_t nbox;            // 128-byte boxes per row (dim * esize / 128)
    uint32_t ring;            // box slots per warp
    uint64_t *out_lists;      // [nq][gridDim.x][k] sorted keys, KEY_SENTINEL padded
    const uint32_t *flag_count;  // flagged mode: device count of queries to run (null = run a.nq queries)
    const uint32_t *flag_list;   // flagged mode: their indices into `queries`
};
size_t scan_smem_bytes(uint32_t W, uint32_t ring, uint32_t qb, uint32_t dim, uint32_t cap);
cudaError_t launch_scan_exact(const CUtensorMap &tmap, const ScanArgs &a, int qb, bool f16, int order, int metric,
                              uint32_t grid, uint32_t W, size_t smem, cudaStream_t st);

// ---- peer push of final results (exchange.cu) ----
// Kernels that write a query's FINAL (ids, dist)[k] can also push it straight into every rank's exchange buffer
// over NVLink peer memory and publish the per-(rank, query) flag, so the multi-GPU step needs no separate
// push / all-gather launch.  world == 0 disables it.  Buffer layout: see exchange.cu.
struct PushArgs {
    uint32_t *peer[8];        // base of every rank's symmetric buffer (peer[rank] is local)
    uint32_t rank, world, epoch, b, k;
    uint32_t q0;              // index of the calling kernel's query 0 within the exchange's b queries
};

// ---- scan_tc.cu (tensor-core sweep) ----
constexpr uint32_t TC_HIST = 32;        // buckets of the per-query score histogram that tightens the threshold
constexpr uint32_t TC_FIN_NC = 2048;    // survivors one finalize CTA can sort (more -> exact re-run)
constexpr uint32_t TC_KK_MAX = 256;     // candidates re-scored per query (k + slack)
struct TcArgs {
    uint64_t n_rows;
    uint32_t dim;
    uint32_t nq;              // live queries in this query tile (128 per tile; 256 in pair mode)
    uint32_t q_tile;          // which query tile
    uint32_t stages;          // B ring depth (stages of kbs K-blocks each)
    uint32_t kk;              // candidates wanted per query (k + slack)
    uint32_t b_pad;           // padded query count: stride of the [value][query] sample array
    uint32_t J;               // sample maxima each region contributes (n_lists * J >= kk, J <= 4)
    uint32_t cap_l;           // entries per (query, list) candidate region
    uint32_t kbs;             // K-blocks per ring stage (1 or 2; divides dim / 64)
    const uint32_t *deleted;
    uint32_t *gtau;           // [b_pad] best threshold key published so far (0xFFFFFFFF = none)
    float *maxima;            // [b_pad][nv_pad] sample maxima (-inf padded); nv_pad = n_lists * J rounded up to 4
    uint32_t nv_pad;
    const float *queries;     // [b][dim] f32: each CTA converts its own 128 rows to the fp16 A operand
    uint32_t qtile;           // queries per sweep (128 / 256)
    uint32_t *arrive;         // CTAs that ever reached the sample meeting point (monotonic, never reset)
    uint32_t arrive_target;   // value of *arrive once every CTA of THIS launch has arrived
    uint32_t *nflag;          // uncertified-query counter of the search (reset by the first sweep of a search)
    uint32_t nv;              // live sample values per query (units * J)
    uint32_t *hist;           // [b_pad][TC_HIST] appended candidates per score bucket
    uint64_t *cand;           // [b_pad][n_lists][cap_l] unsorted approximate candidates
    uint32_t *cnts;           // [b_pad][n_lists] entries appended (may exceed cap_l: overflow)
    unsigned long long *trace; // diagnostics only (SG_TC_TRACE=1): [gridDim.x][8] globaltimer stamps of the epilogue's phases
    float *dump;              // diagnostics only: raw accumulators [dump_rows][nq] (null in production)
    uint64_t dump_row0, dump_rows;
};
struct FinalizeArgs {
    const float *queries;     // [b][dim] f32 (the exact operands)
    const void *rows;         // fp16 corpus
    uint32_t dim, k, kk, order, id_base;
    uint64_t n_live;
    const uint64_t *cand;     // [b][n_lists][cap_l]
    const uint32_t *cnts;     // [b][n_lists]
    uint32_t n_lists, cap_l;
    const uint32_t *gtau;     // [b] final shared threshold key: candidates beyond it cannot be in the top kk
    const float *max_row_norm; // device scalar: upper bound on |row|_2 over the store
    float c_acc;              // accumulation-error coefficient of the bound (derived from dim)
    uint32_t rc;              // candidates re-scored per round (rows staged in shared memory)
    uint32_t *ids_out; float *dist_out; uint32_t *counts_out;
    uint32_t *flags;          // [b] 1 = could not be certified
    uint32_t *n_flagged;      // device counter (reset per search)
    uint32_t *flag_list;      // [b] indices of the uncertified queries, in arrival order
    unsigned long long *flag_total;  // cumulative statistics
    PushArgs push;            // certified queries are pushed to the peers right here
};
size_t tc_smem_bytes(uint32_t dim, uint32_t stages, uint32_t kbs);
float tc_c_acc(uint32_t dim);
cudaError_t launch_scan_tc(const CUtensorMap &tmap_rows, const TcArgs &a, uint32_t grid, size_t smem, cudaStream_t st);
cudaError_t launch_scan_tc2(const CUtensorMap &tmap_rows, const TcArgs &a, uint32_t grid, size_t smem, cudaStream_t st);
cudaError_t launch_finalize_tc(const FinalizeArgs &f, uint32_t b, cudaStream_t st);
// diagnostics: exact -dot of every (query, row) pair in the reference's order -> out [b][n_rows_range]
cudaError_t launch_exact_scores(const float *d_q, uint32_t b, const void *rows, uint32_t dim, uint64_t row0, uint64_t n,
                                int order, float *d_out, cudaStream_t st);

// ---- select.cu ----
// lists [b][n_lists][k] u64 keys (sentinel padded) -> top-k per query.  d_scratch (may be null) holds
// [b][merge_groups(n_lists,k)][k] keys and enables the two-level merge for large n_lists*k.
uint32_t merge_groups(uint32_t n_lists, uint32_t k);
cudaError_t launch_merge_keys(const uint64_t *d_lists, uint32_t b, uint32_t n_lists, uint32_t k, uint32_t id_base,
                              uint64_t *d_scratch, uint32_t *d_ids, float *d_dist, uint32_t *d_counts,
                              cudaStream_t st, const PushArgs *push = nullptr);
// flagged form: CTA i handles slot i of d_lists iff i < *d_slot_count and writes query d_slot_to_query[i]
cudaError_t launch_merge_keys_flagged(const uint64_t *d_lists, uint32_t b, uint32_t n_lists, uint32_t k, uint32_t id_base,
                                      const uint32_t *d_slot_count, const uint32_t *d_slot_to_query, uint32_t *d_ids,
                                      float *d_dist, uint32_t *d_counts, cudaStream_t st, const PushArgs *push = nullptr);
// lists as (ids, dist) pairs [b][n_per_query]; id 0xFFFFFFFF marks an empty slot
cudaError_t launch_merge_pairs(const uint32_t *d_ids_in, const float *d_dist_in, uint32_t b, uint32_t n_lists,
                               uint32_t k, uint64_t list_stride, uint64_t query_stride, uint32_t *d_ids,
                               float *d_dist, uint32_t *d_counts, cudaStream_t st);

// ---- convert / synth (store_kernels.cu) ----
cudaError_t launch_f32_to_f16(const float *d_src, __half *d_dst, uint64_t n_elems, cudaStream_t st);
cudaError_t launch_f16_to_f32(const __half *d_src, float *d_dst, uint64_t n_elems, cudaStream_t st);
cudaError_t launch_synth_rows(void *d_rows, bool f16, uint64_t seed, uint64_t first_row, uint64_t n,
                              uint32_t dim, cudaStream_t st);
cudaError_t launch_set_bit(uint32_t *d_bitmap, uint32_t id, uint32_t *d_was_set, cudaStream_t st);
// max over rows of |row|_2 (rounded up), folded into *d_max_norm (a float, initialised to 0)
cudaError_t launch_max_row_norm(const void *d_rows, bool f16, uint64_t n, uint32_t dim, float *d_max_norm,
                                cudaStream_t st);

// ---- fuse.cu ----
struct FuseDev {
    const float *semantic, *entity, *tag, *importance, *momentum, *graph;
    const uint32_t *access;
    const int64_t *created;
};
cudaError_t launch_fuse_rank(const FuseDev &f, uint32_t n, sg_fuse_weights w, sg_fuse_config c,
                             int64_t now_s, uint32_t *d_order, float *d_score, uint32_t *d_n_out,
                             float *d_all, cudaStream_t st);

// ---- post.cu ----
cudaError_t launch_rrf_fuse(float k, const float *d_weights, uint32_t n_weights, const uint64_t *d_ids, const uint32_t *d_offsets,
                            uint32_t n_lists, uint32_t n_total, uint64_t *d_ids_out, float *d_score_out, uint32_t *d_n_out,
                            cudaStream_t st);
cudaError_t launch_search_ids_post(const uint32_t *d_ids, const float *d_dist, const uint32_t *d_counts, uint32_t b, uint32_t k,
                                   const uint64_t *d_vec2mem, uint64_t n_map, uint32_t id_base, uint32_t limit,
                                   uint64_t *d_mem_out, float *d_sim_out, uint32_t *d_cnt_out, cudaStream_t st);

// ---- vamana.cu (graph index: beam search, incremental linking, medoid, bulk build) ----
struct AnnArgs {
    const void *rows;            // the store's row matrix
    uint32_t dim;
    uint64_t n_rows;             // rows in storage (ids >= n_rows in a list are ignored)
    const uint32_t *graph;       // [graph_len][gs] neighbour ids, 0xFFFFFFFF-padded
    uint32_t gs;                 // list stride = max_degree
    uint64_t graph_len;          // graph.len(): nodes beyond it have no list yet (vamana.rs:606-609)
    uint32_t entry;              // medoid
    const float *queries;        // [nq][dim] f32, or null when query_rows is set
    const uint32_t *query_rows;  // query q = store row query_rows[q] (build / add_vector)
    uint32_t nq;
    uint32_t K;                  // beam width = the k greedy_search is called with
    uint32_t k_out;              // results written per query (after dropping deleted ids)
    const uint32_t *deleted;     // soft-delete bitmap or null
    uint32_t id_base;
    uint32_t *vis;               // [nq][H] visited hash tables, 0xFFFFFFFF-filled
    uint32_t H;                  // power of two
    uint32_t *ids_out; float *dist_out; uint32_t *counts_out;   // [nq][k_out], [nq][k_out], [nq]
    uint32_t *flags;             // [nq] or null: bit 0 = visited table filled up (result incomplete), bit 1 = > 32 ties dropped
    unsigned long long *stats;   // [nq][2] or null: nodes expanded, distance evaluations
    unsigned long long *phase;   // [8] or null: SM cycles per phase summed over queries (profiling hook)
};
// shared memory per query warp: query f32 + two copies of the sorted list (K + 32 keys + expanded flags each:
// the batched merge writes the new list beside the old one) + 32 candidate keys + the visited set's 1 KB slot arbiter
__host__ __device__ inline size_t ann_warp_smem_bytes(uint32_t dim, uint32_t K)
{
    const size_t cap = (size_t)K + 32;   // every piece a multiple of 16 bytes: the query is read as float4
    return (((size_t)dim * 4 + 15) & ~(size_t)15) + 2 * (((cap * 8 + 15) & ~(size_t)15) + ((cap + 15) & ~(size_t)15)) + 256 + 1024;
}
// staging buffer of the beam search: 32 rows, each padded by 16 bytes (conflict-free lane-per-row reads)
__host__ __device__ inline size_t ann_stage_smem_bytes(uint32_t dim, bool f16)
{
    return (size_t)32 * ((size_t)dim * (f16 ? 2 : 4) + 16) + 16;   // + the warp's mbarrier
}
cudaError_t launch_ann_search(const AnnArgs &a, bool f16, int order, int metric, cudaStream_t st);
cudaError_t launch_graph_link(const void *rows, bool f16, int order, int metric, uint32_t dim, uint32_t *graph, uint32_t gs,
                              uint32_t id, const uint32_t *res_ids, const uint32_t *res_count, cudaStream_t st);
cudaError_t launch_centroid(const void *rows, bool f16, uint32_t dim, uint64_t n, float *d_centroid, cudaStream_t st);
cudaError_t launch_graph_init(uint32_t *graph, uint32_t gs, uint64_t n, uint64_t seed, cudaStream_t st);
cudaError_t launch_iota(uint32_t *out, uint32_t first, uint32_t n, cudaStream_t st);
cudaError_t launch_graph_degree(const uint32_t *graph, uint32_t gs, uint64_t n, uint32_t *deg, cudaStream_t st);
cudaError_t launch_build_prune(const void *rows, bool f16, uint32_t dim, int metric, float alpha, const uint32_t *graph,
                               uint32_t gs, const uint32_t *node_ids, uint32_t n_nodes, const uint32_t *res_ids,
                               const uint32_t *res_counts, uint32_t L, uint32_t *new_lists, cudaStream_t st);
cudaError_t launch_build_apply(uint32_t *graph, uint32_t gs, const uint32_t *node_ids, uint32_t n_nodes,
                               const uint32_t *new_lists, uint32_t *inbox, uint32_t *inbox_cnt, uint32_t ecap,
                               uint32_t *touched, uint32_t *n_touched, cudaStream_t st);
cudaError_t launch_build_backedges(const void *rows, bool f16, uint32_t dim, int metric, float alpha, uint32_t *graph,
                                   uint32_t gs, const uint32_t *touched, const uint32_t *n_touched, const uint32_t *inbox,
                                   uint32_t *inbox_cnt, uint32_t ecap, bool by_id, uint32_t grid, cudaStream_t st);

// ---- tensor map (store.cu) ----
int encode_corpus_tmap(CUtensorMap *out, const void *d_rows, bool f16, uint32_t dim, uint64_t n_rows,
                       uint32_t box_rows);

}  // namespace sg
