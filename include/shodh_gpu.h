/*
 * shodh_gpu.h -- C ABI of libshodh_gpu.so: the B200-native (sm_100a) drop-in for
 * the semantic-recall hot path of shodh-memory (query embedding -> exact
 * cosine/dot top-k scan over the in-process vector store -> relevance fusion).
 *
 * The reference has no FFI for this path; its seams are Rust-level (SURVEY.md
 * section 8b).  Each entry point below names the reference interface it
 * replaces (paths relative to the reference repo).  INTEGRATION.md shows the
 * Rust `extern "C"` block and the `GpuVectorStore` / `GpuMiniLM` wrappers a
 * maintainer would add behind `VamanaIndex` / `Embedder`.
 *
 * Conventions (all functions):
 *   - return SG_OK (0) or a negative sg_status; never throw, never abort;
 *   - NO CPU fallback: if there is no usable sm_100 device the call fails;
 *   - caller owns every host buffer; handles own all device memory;
 *   - a handle may be used from many threads; calls on one handle serialise;
 *   - sg_last_error() returns a thread-local, human-readable message.
 */
#ifndef SHODH_GPU_H
#define SHODH_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SG_API __attribute__((visibility("default")))
#else
#define SG_API
#endif

typedef enum sg_status {
    SG_OK = 0,
    SG_ERR_INVALID = -1,      /* bad argument (null pointer, k == 0, id out of range ...) */
    SG_ERR_CUDA = -2,         /* a CUDA runtime/driver call failed; see sg_last_error() */
    SG_ERR_OOM = -3,          /* device or pinned-host allocation failed */
    SG_ERR_DIM = -4,          /* dimension mismatch (spann.rs:586-592 sets the precedent: error, not UB) */
    SG_ERR_UNSUPPORTED = -5,  /* valid request outside what the kernels cover (documented limits) */
    SG_ERR_NO_DEVICE = -6,    /* no sm_100 device / extension cannot run: fail loudly, never fall back */
    SG_ERR_NOT_BUILT = -7     /* store holds no vectors where the reference would say "graph not built" */
} sg_status;

/* element type the corpus is held in on the device (queries are always f32) */
typedef enum sg_dtype { SG_DTYPE_F32 = 0, SG_DTYPE_F16 = 1 } sg_dtype;

/* src/vector_db/vamana.rs:38-54 DistanceMetric, dispatched by VamanaIndex::distance (:755-761).
 * RetrievalEngine only permits NormalizedDotProduct (src/memory/retrieval.rs:188-193); the other
 * two are what a standalone VamanaIndex accepts.  All are scored in the reference's summation order
 * by the exact sweep; only NEG_DOT batches use the tensor-core sweep.
 * NaN distances: every NaN is canonicalised to +NaN and sorts LAST (f32::total_cmp order of a positive NaN).
 * For NEG_DOT that is what x86 produces for the reference (default NaN from arithmetic, negated by `-dot`...
 * canonicalised the same way in the oracle); for EUCLIDEAN / COSINE the reference's NaN sign follows x86 NaN
 * propagation through sub/mul/sqrt/div and is NOT reproduced -- rows whose distance is NaN are still returned
 * after all numbered distances, but their relative order among themselves is by id only. */
typedef enum sg_metric {
    SG_METRIC_NEG_DOT = 0,    /* normalized_distance_inline (distance_inline.rs:479-481): -dot */
    SG_METRIC_EUCLIDEAN = 1,  /* euclidean_squared_inline (:184-305): |a-b|^2 */
    SG_METRIC_COSINE = 2,     /* 1 - cosine_similarity_inline (:444-458) */
    SG_METRIC_COSINE_SIM = 3  /* -cosine_similarity of src/similarity.rs:10-24 ranked as top_k_similar does
                                 (:27-48: score desc, ties keep input order); what sg_top_k_similar runs on */
} sg_metric;

/* which summation order the exact scores reproduce */
typedef enum sg_order {
    SG_ORDER_SCALAR = 0, /* distance_inline.rs:157-173, what a default x86-64 build runs */
    SG_ORDER_AVX2 = 1    /* distance_inline.rs:67-111, `-C target-cpu=native` builds */
} sg_order;

typedef struct sg_store sg_store;
typedef struct sg_encoder sg_encoder;

/* ---- library ------------------------------------------------------------------------- */
SG_API const char *sg_version(void);
SG_API const char *sg_last_error(void);
SG_API int sg_device_count(int *count_out);
/* number of this library's kernels launched since load (bench.py: gpu_launches) */
SG_API uint64_t sg_kernel_launch_count(void);

/* ---- vector store: replaces VamanaIndex as held by RetrievalEngine ---------------------
 * (src/memory/retrieval.rs:53 `vector_index: Arc<RwLock<VamanaIndex>>`).
 * ids are dense and sequential from 0 in insertion order (vamana.rs:854-855). */

/* VamanaIndex::new (vamana.rs:163-198).  dim must be a multiple of 64 (F16) / 32 (F32),
 * <= 2048.  device = CUDA ordinal. */
SG_API int sg_store_create(sg_store **out, int device, uint32_t dim, int dtype, int metric);
SG_API void sg_store_destroy(sg_store *s);

/* pre-size device memory for n_rows (amortises add_vector growth) */
SG_API int sg_store_reserve(sg_store *s, uint64_t n_rows);
/* How the row matrix is held.  in_place_out = 1: it lives in a reserved virtual range and grows by mapping physical
 * chunks at its end (no second copy of the corpus while growing, no device-to-device copy, base address fixed);
 * 0: the allocate-x1.5-and-copy fallback (driver without virtual memory management, or SG_STORE_VMM=0).
 * capacity_rows_out = rows that fit without growing, bytes_out = device bytes held for rows, chunks_out = mapped chunks. */
SG_API int sg_store_memory_info(sg_store *s, int *in_place_out, uint64_t *capacity_rows_out, uint64_t *bytes_out,
                                uint32_t *chunks_out);

/* VamanaIndex::add_vector (vamana.rs:853-974) / build (:200-284) /
 * rebuild_from_vectors: append n row-major f32 rows (rounded to the store dtype
 * on the device, round-to-nearest-even).  first_id_out receives the id of row 0. */
SG_API int sg_store_append(sg_store *s, const float *rows_f32, uint64_t n, uint32_t *first_id_out);
/* the same for rows that are already on the device (e.g. sg_encoder_encode_device output: batched document
 * ingest, retrieval.rs:668-676 -> minilm.rs:996, without a host round trip); returns after the rows are in. */
SG_API int sg_store_append_device(sg_store *s, const float *d_rows_f32, uint64_t n, uint32_t *first_id_out,
                                  void *cuda_stream);

/* VamanaIndex::mark_deleted (vamana.rs:813-820): soft delete.  *ok_out = 1 if id < len
 * (the reference returns true, also for an id that was already deleted), 0 otherwise. */
SG_API int sg_store_mark_deleted(sg_store *s, uint32_t id, int *ok_out);
/* VamanaIndex::is_deleted (:823-825) / clear_deleted (:850-852) */
SG_API int sg_store_is_deleted(sg_store *s, uint32_t id, int *deleted_out);
SG_API int sg_store_clear_deleted(sg_store *s);

/* VamanaIndex::len (vamana.rs:976-979): rows ever appended (deleted included). */
SG_API int sg_store_len(const sg_store *s, uint64_t *n_out);
SG_API int sg_store_deleted_count(const sg_store *s, uint64_t *n_out);
SG_API int sg_store_dim(const sg_store *s, uint32_t *dim_out);
SG_API int sg_store_metric(const sg_store *s, int *metric_out);

/* VamanaIndex::get_vector / extract_all_vectors: rows [first, first+n) upcast to f32. */
SG_API int sg_store_get_vectors(sg_store *s, uint64_t first, uint64_t n, float *rows_f32_out);

/* VamanaIndex::load_from_file (src/vector_db/vamana_persist.rs:290-400): read the reference's persisted
 * "VAMA" v1 index file into an EMPTY store -- vectors section (dense f32 rows) uploaded as-is, deleted ids
 * applied, graph section + medoid uploaded (sg_store_set_graph) so sg_store_search_ann walks the graph the
 * reference built; the exact sweep ignores it.  The store's dim and metric must equal the file's.
 * verify_checksum != 0 checks the FNV-1a-64 checksum like the reference does (:303-312). */
SG_API int sg_store_load_vama(sg_store *s, const char *path, int verify_checksum);
/* VamanaIndex::save_to_file (vamana_persist.rs:175-280): the store (rows upcast to f32, deleted ids, graph and
 * medoid if it has them, else empty lists) written as a VAMA v1 file the reference's load_from_file accepts. */
SG_API int sg_store_save_vama(sg_store *s, const char *path);
/* header / layout of a VAMA file without touching a GPU (host-side parser only) */
SG_API int sg_vama_inspect(const char *path, int verify_checksum, uint64_t *num_vectors_out, uint32_t *dimension_out,
                           uint32_t *deleted_count_out, uint32_t *metric_out, uint64_t *vectors_offset_out);

/* summation order of the exact scores (default SG_ORDER_SCALAR) */
SG_API int sg_store_set_order(sg_store *s, int order);

/* row-sharding: ids reported by search are id_base + local id */
SG_API int sg_store_set_id_base(sg_store *s, uint32_t id_base);

/* VamanaIndex::search (vamana.rs:764-808) on its exact path (`SHODH_VECTOR_EXACT`,
 * :775-777 -> brute_force_search :1167-1188), batched over b queries:
 * for every live row dist = -dot(query,row) in the reference's summation order,
 * results sorted by (dist total_cmp ascending, id ascending), truncated to k.
 * queries_f32 [b,dim]; ids_out/dist_out [b,k]; counts_out [b] = min(k, live rows).
 * Unused tail slots hold id 0xFFFFFFFF / dist +inf.  Empty store -> counts 0 (Ok(vec![])).
 * Limit: k <= 1024. */
SG_API int sg_store_search(sg_store *s, const float *queries_f32, uint32_t b, uint32_t k,
                           uint32_t *ids_out, float *dist_out, uint32_t *counts_out);

/* src/similarity.rs:27-48 top_k_similar(query, candidates, k) with cosine_similarity (:10-24: three dots in
 * dot_product_inline's order, 0 on a zero norm or length mismatch, clamp to [-1,1]): candidates is a dense
 * [n,dim] f32 matrix (the reference's &[(Vec<f32>,T)] with T = the row position), scored on the device in the
 * reference's summation order and ranked (score desc, position asc) -- OrderedFloat order, stable sort.
 * Writes min(k,n) positions and scores.  dim must be a multiple of 32, k <= 1024. */
SG_API int sg_top_k_similar(int device, const float *query, uint32_t query_dim, const float *candidates, uint64_t n,
                            uint32_t dim, uint32_t k, int order, uint32_t *pos_out, float *score_out, uint32_t *n_out);

/* RetrievalEngine::search_ids from the embedding on (src/memory/retrieval.rs:905-964): the vector id ->
 * MemoryId table of IdMapping (vec2mem[i] = the Uuid of the memory that owns vector i as two big-endian
 * u64 halves; all-ones = no mapping) is registered once; a search then asks the scan for
 * limit * VECTOR_SEARCH_CANDIDATE_MULTIPLIER(2) * 2 candidates (:915-916), and ON THE DEVICE converts
 * similarity = -distance, keeps the best chunk per memory, sorts (similarity desc, MemoryId asc) and
 * truncates to limit (:927-961).  mem_out [b,limit,2] u64, sim_out [b,limit], counts_out [b]; limit <= 256. */
SG_API int sg_store_set_memory_ids(sg_store *s, const uint64_t *vec2mem, uint64_t n);
SG_API int sg_store_search_memory_ids(sg_store *s, const float *queries_f32, uint32_t b, uint32_t limit,
                                      uint64_t *mem_out, float *sim_out, uint32_t *counts_out);

/* same, with queries and results resident in device memory and asynchronous on
 * `cuda_stream` (a cudaStream_t; NULL = the CUDA default stream).  This is the
 * form the GPU encoder feeds and the multi-GPU exchange consumes.
 * Streams: a store owns ONE set of scratch buffers.  Every call on a store -- asynchronous or not, on any
 * stream -- first waits (cudaStreamWaitEvent) for the end of the previous asynchronous call if that ran on a
 * different stream, so calls on one store are serialised on the device in the order they were issued.  Results
 * of an asynchronous call are valid once `cuda_stream` has passed the call. */
SG_API int sg_store_search_device(sg_store *s, const float *d_queries_f32, uint32_t b, uint32_t k,
                                  uint32_t *d_ids_out, float *d_dist_out, uint32_t *d_counts_out,
                                  void *cuda_stream);

/* Batched searches (b >= 3, k <= 128 -- which covers the reference's own request, k = 4 * limit = 120 for a
 * top-10 recall: retrieval.rs:912-918 x memory/mod.rs:3585 --, F16 store, NormalizedDotProduct) run the
 * tensor-core sweep: approximate scores select k+slack candidates, which are re-scored exactly in the
 * reference's summation order and certified with a rigorous error bound (coefficient derived from dim);
 * queries that cannot be certified are re-run on the exact sweep, so results are always the exact ones.
 * Statistics: queries that took the tensor path / that needed the re-run. */
SG_API int sg_store_tc_stats(sg_store *s, uint64_t *queries_out, uint64_t *flagged_out);
/* mode 0 (default): choose the sweep by shape; mode 1: always the exact sweep (path A) -- what the parity
 * checks in tests/ and bench.py compare the tensor-core path against. */
SG_API int sg_store_set_tc_mode(sg_store *s, int mode);
/* Diagnostics for tests/test_scan_tc_gpu.py (not part of the drop-in surface): raw tensor-core scores and
 * exact reference-order dots of b queries against rows [first_row, first_row + n_rows), so the accumulation-error
 * coefficient of the certificate can be MEASURED.  sg_debug_tc_scores writes out[n_rows][b] (row-major by corpus
 * row: the layout the sweep's epilogue can store coalesced), sg_debug_exact_scores out[b][n_rows].  F16 stores;
 * b <= 256 and dim <= 704 for the tensor-core form. */
SG_API int sg_debug_tc_scores(sg_store *s, const float *queries_f32, uint32_t b, uint64_t first_row, uint64_t n_rows,
                              float *out);
SG_API int sg_debug_exact_scores(sg_store *s, const float *queries_f32, uint32_t b, uint64_t first_row,
                                 uint64_t n_rows, float *out);

/* k-way merge of per-shard results: lists [b, n_lists, k] of (id, dist)
 * already sorted by (dist, id) -> the global top-k per query, same order rule.
 * All pointers are device pointers; asynchronous on cuda_stream. */
SG_API int sg_merge_topk_device(int device, const uint32_t *d_ids, const float *d_dist,
                                uint32_t b, uint32_t n_lists, uint32_t k, uint32_t *d_ids_out,
                                float *d_dist_out, uint32_t *d_counts_out, void *cuda_stream);

/* general layout: element (query q, list l, slot j) lives at l*list_stride + q*query_stride + j
 * in both d_ids and d_dist (e.g. the receive buffer of an all-gather: list_stride = per-rank block). */
SG_API int sg_merge_topk_device_strided(int device, const uint32_t *d_ids, const float *d_dist,
                                        uint32_t b, uint32_t n_lists, uint32_t k, uint64_t list_stride,
                                        uint64_t query_stride, uint32_t *d_ids_out, float *d_dist_out,
                                        uint32_t *d_counts_out, void *cuda_stream);

/* Fused exchange + merge over NVLink peer memory (the product path for N > 1; the NCCL all-gather +
 * sg_merge_topk_device_strided pair is the baseline it replaces).  peer_buffers[r] = device address of
 * rank r's symmetric buffer of sg_exchange_buffer_words(world,b,k) 32-bit words, zero-initialised once and
 * mapped on every rank (e.g. torch symmetric memory).  Every rank calls this with the same (b, k) and an
 * epoch that increases by 1 per call (>= 1): the kernel pushes this rank's lists into every peer, signals
 * per (rank, query) flags, waits for the peers' flags and merges -- no collective launch, no barrier. */
SG_API uint64_t sg_exchange_buffer_words(uint32_t world, uint32_t b, uint32_t k);
SG_API int sg_exchange_merge_topk_device(int device, const uint64_t *peer_buffers, uint32_t rank, uint32_t world,
                                         uint32_t epoch, const uint32_t *d_ids, const float *d_dist, uint32_t b,
                                         uint32_t k, uint32_t *d_ids_out, float *d_dist_out, uint32_t *d_counts_out,
                                         void *cuda_stream);
/* The whole multi-GPU recall step of one rank: search this shard (sg_store_search_device) and
 * exchange + merge, with the push folded into the kernels that produce each query's final local result
 * (the re-score/certify kernel of the tensor-core path, the merge kernel of the exact path): a query's k
 * results travel to every peer the moment they exist, and the only extra launch is the wait + merge kernel.
 * Same buffers, epoch rule and (b, k) agreement as sg_exchange_merge_topk_device.  Outputs: the GLOBAL top-k. */
SG_API int sg_store_search_sharded_device(sg_store *s, const float *d_queries_f32, uint32_t b, uint32_t k,
                                          const uint64_t *peer_buffers, uint32_t rank, uint32_t world, uint32_t epoch,
                                          uint32_t *d_ids_out, float *d_dist_out, uint32_t *d_counts_out,
                                          void *cuda_stream);

/* ---- Vamana graph index: the reference's DEFAULT search arm -------------------------
 * src/vector_db/vamana.rs: build :200-284, search :779-807 -> greedy_search :576-657, add_vector :853-973.
 * The exact sweep above needs no graph; these entry points exist so that an index the reference built (or
 * its VAMA file, sg_store_load_vama) is searched the way the reference searches it, and so that corpora
 * beyond the exact sweep's reach (>= 10^8 rows per GPU) keep a sub-millisecond path.
 *
 * sg_store_search_ann: VamanaIndex::search without SHODH_VECTOR_EXACT: beam search from the medoid with
 * beam k (+ min(deleted, 2k) when rows are soft-deleted, :785-796), deleted ids dropped, k results
 * (id, distance) ascending.  GIVEN THE SAME GRAPH THE RESULT IS THE REFERENCE'S, id for id and bit for bit:
 * SearchCandidate's order (distance total_cmp, id) is total, so the heaps' pops are determined, and every
 * distance is evaluated in the reference's summation order.  SG_ERR_NOT_BUILT when rows exist but no graph
 * ("Vamana graph not built", :780-784).  stats_out (may be NULL) [b,2]: nodes expanded, distance
 * evaluations per query.  Limits: k <= 1024, beam <= 2048, max_degree <= 64.
 *
 * sg_store_build_graph: VamanaIndex::build over the rows already in the store: random start graph, medoid
 * (centroid summed in row order, arg-min by the exact sweep -- the reference's medoid exactly), two passes of
 * beam search (search_list_size) + robust_prune (alpha-RNG rule with the +1 offset for NormalizedDotProduct,
 * :711-727) + back-edges.  Nodes are processed in batches, not one by one, so the graph differs from a
 * sequential build edge for edge (the reference's own start graph is random per run); the contract is
 * recall against the exact sweep (tests/test_vamana_gpu.py).  Defaults of VamanaConfig: 32, 75, 1.2.
 *
 * sg_store_set_graph / sg_store_get_graph: upload / download the adjacency (deg [n], nbr [n,stride]).
 * sg_store_enable_graph: an EMPTY store grows its graph incrementally from the first sg_store_append on,
 * one VamanaIndex::add_vector per row (beam max_degree, neighbours as found, back-edges kept to the
 * max_degree closest by (distance, id), :905-966) -- edge for edge the reference's sequence.  Once a store
 * has a graph (built, set, loaded or enabled), every later append links its rows the same way. */
SG_API int sg_store_enable_graph(sg_store *s, uint32_t max_degree, uint32_t search_list_size, float alpha);
SG_API int sg_store_build_graph(sg_store *s, uint32_t max_degree, uint32_t search_list_size, float alpha, uint64_t seed);
SG_API int sg_store_set_graph(sg_store *s, uint32_t max_degree, uint32_t medoid, const uint32_t *deg, const uint32_t *nbr,
                              uint32_t stride, uint64_t n_nodes);
SG_API int sg_store_get_graph(sg_store *s, uint32_t *deg_out, uint32_t *nbr_out /* [n_nodes, max_degree] */,
                              uint64_t n_nodes);
SG_API int sg_store_graph_info(sg_store *s, uint32_t *max_degree_out, uint32_t *medoid_out, uint64_t *graph_len_out);
SG_API int sg_store_search_ann(sg_store *s, const float *queries_f32, uint32_t b, uint32_t k, uint32_t *ids_out,
                               float *dist_out, uint32_t *counts_out, uint64_t *stats_out);
/* device pointers, asynchronous on cuda_stream; d_flags (may be NULL) [b]: bit 0 = the query's visited table
 * filled up and its result is incomplete (the host-buffer form re-runs such queries with a larger table) */
SG_API int sg_store_search_ann_device(sg_store *s, const float *d_queries_f32, uint32_t b, uint32_t k, uint32_t *d_ids_out,
                                      float *d_dist_out, uint32_t *d_counts_out, uint32_t *d_flags, void *cuda_stream);
SG_API int sg_store_ann_stats(sg_store *s, uint64_t *queries_out, uint64_t *retries_out);
/* profiling hook: cycles_out[0..5) (may be NULL) = SM cycles the beam-search warps spent, summed over all queries
 * since the last call, in: pop, list read + visited set, row fetch, scoring, insertion.  enable != 0 keeps
 * counting (a clock64 read per phase), 0 switches it off. */
SG_API int sg_store_ann_phase_cycles(sg_store *s, int enable, uint64_t *cycles_out);

/* ---- Product Quantization: src/vector_db/pq.rs ----------------------------------------
 * ProductQuantizer (train :94-215, encode :220-257, decode :260-290, build_distance_table :329-353,
 * distance_with_table :358-369) and CompressedVectorStore (add :428-432, search :449-466) on the device, every
 * f32 operation in the reference's order, so codes, tables, distances and the (distance, id) ranking are
 * bit-identical to the CPU path.  centroids: [dim / subvec_dim][num_centroids][subvec_dim] f32.
 *   sg_pq_create  a quantizer from trained centroids (the reference's struct is serde-serialisable);
 *   sg_pq_train   k-means per subspace like ProductQuantizer::train; `orders` [num_subvectors][n] are the shuffled
 *                 index orders the reference draws from thread_rng inside kmeans (:157-159) -- given the same
 *                 orders the centroids equal the CPU's bit for bit; num_centroids is capped at n (:118);
 *   sg_pq_add     CompressedVectorStore::add for n vectors (ids must be unique: the reference's map would overwrite);
 *   sg_pq_search  CompressedVectorStore::search for b queries: ids_out / dist_out [b, k], counts_out [b].
 * Limits: subvec_dim <= 16, dim / subvec_dim a multiple of 4, num_centroids <= 256, k <= 1024. */
typedef struct sg_pq sg_pq;
SG_API int sg_pq_create(sg_pq **out, int device, uint32_t dim, uint32_t subvec_dim, uint32_t num_centroids,
                        const float *centroids);
SG_API int sg_pq_train(sg_pq **out, int device, uint32_t dim, uint32_t subvec_dim, uint32_t num_centroids,
                       uint32_t iterations, const float *vectors_f32, uint64_t n, const uint32_t *orders);
SG_API void sg_pq_destroy(sg_pq *p);
SG_API int sg_pq_info(sg_pq *p, uint32_t *dim_out, uint32_t *num_subvectors_out, uint32_t *num_centroids_out,
                      uint64_t *len_out);
SG_API int sg_pq_get_centroids(sg_pq *p, float *centroids_out);
SG_API int sg_pq_encode(sg_pq *p, const float *rows_f32, uint64_t n, uint8_t *codes_out /* [n, num_subvectors] */);
SG_API int sg_pq_decode(sg_pq *p, const uint8_t *codes, uint64_t n, float *rows_f32_out);
SG_API int sg_pq_table(sg_pq *p, const float *query_f32, float *table_out /* [num_subvectors, num_centroids] */);
SG_API int sg_pq_add(sg_pq *p, const uint32_t *ids, const float *rows_f32, uint64_t n);
SG_API int sg_pq_search(sg_pq *p, const float *queries_f32, uint32_t b, uint32_t k, uint32_t *ids_out, float *dist_out,
                        uint32_t *counts_out);

/* ---- SPANN partitioned index: src/vector_db/spann.rs ---------------------------------
 * SpannIndex::{build :363-463, search :574-693, insert :1006-1051}: P = ceil(sqrt(n)) k-means partitions over the
 * full vectors, a product quantizer, per partition a posting list of (vector id, PQ code).  A search ranks the
 * partitions by compute_distance(query, centroid) (:562-572; ties by index), probes the first num_probes and keeps
 * the k smallest ADC distances, sorted (distance total_cmp, id).  One CTA per query does all of it; every f32
 * operation is the reference's, so probes, distances and ids match the CPU path bit for bit.
 *   sg_spann_build   `order` [n] and `pq_orders` [dim/8][n] are the shuffled index orders the reference draws from
 *                    thread_rng inside its two k-means; num_partitions = 0 -> ceil(sqrt(n)); PQ = PQConfig::for_dimension.
 *   sg_spann_create  from arrays (an index built elsewhere): centroids [P,dim], PQ centroids [dim/ds][kc][ds],
 *                    posting lists as CSR part_off [P+1] + ids / codes grouped by partition.
 *   sg_spann_get     the same arrays back (any pointer may be NULL).
 * metric: 0 NormalizedDotProduct / 2 Cosine -> 1 - dot, 1 Euclidean.  Limits: P <= 4096, k <= 512. */
typedef struct sg_spann sg_spann;
SG_API int sg_spann_build(sg_spann **out, int device, uint32_t dim, uint32_t num_partitions, uint32_t num_probes,
                          uint32_t kmeans_iterations, int metric, const float *vectors_f32, uint64_t n,
                          const uint32_t *order, uint32_t pq_iterations, const uint32_t *pq_orders);
SG_API int sg_spann_create(sg_spann **out, int device, uint32_t dim, uint32_t num_partitions, uint32_t num_probes, int metric,
                           const float *centroids, const float *pq_centroids, uint32_t subvec_dim, uint32_t num_centroids,
                           const uint64_t *part_off, const uint32_t *ids, const uint8_t *codes);
SG_API void sg_spann_destroy(sg_spann *s);
SG_API int sg_spann_info(sg_spann *s, uint32_t *num_partitions_out, uint64_t *len_out, uint32_t *num_probes_out);
SG_API int sg_spann_get(sg_spann *s, float *centroids_out, float *pq_centroids_out, uint64_t *part_off_out,
                        uint32_t *ids_out, uint8_t *codes_out);
SG_API int sg_spann_shape(sg_spann *s, uint32_t *dim_out, uint32_t *num_subvectors_out, uint32_t *num_centroids_out,
                          uint32_t *subvec_dim_out, int *metric_out);
SG_API int sg_spann_set_probes(sg_spann *s, uint32_t num_probes);
/* SpannIndex::save_to_file / load_from_file / verify_index_file (spann.rs:750-1003, :1069-1083): the reference's
 * "SPAN" v1 index file (128-byte header, centroids, PQ codebook, posting index, posting data; FNV-1a-64 checksum).
 * The sg_spann_file_* functions are plain host code (no device): inspect = header + optional checksum check,
 * read = the arrays sg_spann_create takes (sizes from inspect), write = the same arrays to a file the reference
 * loads.  load_file / save_file compose them with sg_spann_create / sg_spann_get. */
SG_API int sg_spann_file_inspect(const char *path, int verify_checksum, uint64_t *num_vectors_out, uint32_t *num_partitions_out,
                                 uint32_t *dimension_out, uint32_t *num_subvectors_out, uint32_t *num_centroids_out,
                                 uint32_t *subvec_dim_out, uint32_t *metric_out);
SG_API int sg_spann_file_read(const char *path, float *centroids_out, float *pq_centroids_out, uint64_t *part_off_out,
                              uint32_t *ids_out, uint8_t *codes_out);
SG_API int sg_spann_file_write(const char *path, uint32_t dim, uint32_t num_partitions, uint32_t metric, const float *centroids,
                               uint32_t num_subvectors, uint32_t num_centroids, uint32_t subvec_dim, const float *pq_centroids,
                               const uint64_t *part_off, const uint32_t *ids, const uint8_t *codes);
SG_API int sg_spann_load_file(sg_spann **out, int device, const char *path, uint32_t num_probes, int verify_checksum);
SG_API int sg_spann_save_file(sg_spann *s, const char *path);
SG_API int sg_spann_insert(sg_spann *s, uint32_t vector_id, const float *vector_f32);
SG_API int sg_spann_search(sg_spann *s, const float *queries_f32, uint32_t b, uint32_t k, uint32_t *ids_out,
                           float *dist_out, uint32_t *counts_out);

/* CUDA-event timing of the scan kernel(s) inside each search (bench.py roofline): with
 * profiling on, every search records an event pair around its sweep launches on the stream
 * it runs on (no synchronisation).  sg_store_profile_read waits for the last recorded
 * search, writes up to cap per-search durations (ms) in call order, returns how many
 * searches were recorded since the last read/reset and the sweep launches per search. */
SG_API int sg_store_set_profiling(sg_store *s, int on);
SG_API int sg_store_profile_read(sg_store *s, float *ms_out, uint32_t cap, uint32_t *n_out,
                                 uint32_t *launches_per_search_out);
SG_API int sg_store_last_scan_ms(sg_store *s, float *ms_out, uint32_t *launches_out);

/* TEST/BENCH ONLY: fill rows [first_row, first_row+n) of the deterministic synthetic
 * corpus (oracle/shodh_oracle.c so_synth_corpus, bit-identical) directly on the device. */
SG_API int sg_store_append_synthetic(sg_store *s, uint64_t seed, uint64_t first_row, uint64_t n);

/* ---- relevance fusion: replaces the ranking phase of
 * RelevanceEngine::surface_relevant_inner (src/relevance.rs:805-918) -------------------- */
typedef struct sg_fuse_weights { /* LearnedWeights, relevance.rs:343-369; defaults :64-88 */
    float semantic, entity, tag, importance, momentum, access_count, graph_strength;
} sg_fuse_weights;

typedef struct sg_fuse_config { /* RelevanceConfig, relevance.rs:120-195 */
    uint64_t recency_boost_hours;   /* 24   */
    float recency_boost_multiplier; /* 1.2  */
    float graph_boost_multiplier;   /* 1.15 */
    float min_score;                /* 0.25 (MIN_RELEVANCE_SCORE, relevance.rs:914) */
    uint32_t max_results;           /* 5    */
} sg_fuse_config;

typedef struct sg_candidate_feats { /* SoA over n candidates, host pointers */
    const float *semantic;        /* similarity = -distance from the scan */
    const float *entity;
    const float *tag;
    const float *importance;
    const float *momentum;        /* feedback EMA in [-1,1] */
    const uint32_t *access_count;
    const float *graph_strength;
    const int64_t *created_at_unix_s;
} sg_candidate_feats;

SG_API void sg_fuse_default_weights(sg_fuse_weights *w);
SG_API void sg_fuse_default_config(sg_fuse_config *c);

/* fuse_scores_full (:529-594) + apply_recency_boost (:1524-1547) + graph boost
 * (:877-881) + sort (score desc, created_at desc, candidate position asc) (:905-910)
 * + `>= min_score` (:914-915) + truncate (:918).  order_out/score_out [max_results];
 * all_scores_out [n] optional (may be NULL). */
SG_API int sg_fuse_topk(int device, const sg_fuse_weights *w, const sg_fuse_config *cfg,
                        const sg_candidate_feats *feats, uint32_t n, int64_t now_unix_s,
                        uint32_t *order_out, float *score_out, uint32_t *n_out,
                        float *all_scores_out);

/* RRFusion::new(k, weights).fuse(ranked_lists) (src/memory/hybrid_search.rs:537-594): the hand-off from the
 * retrieval legs (BM25 list from the host, vector list from sg_store_search_memory_ids) to the fused
 * candidate list.  mem_ids [n_total,2] u64 = the lists concatenated in rank order (Uuid as two big-endian
 * halves); list_offsets [n_lists+1]; weights [n_weights] RAW (normalised here exactly like RRFusion::new;
 * lists beyond n_weights use 1.0 like `weights.get(i).unwrap_or(1.0)`).  Output: the distinct ids with
 * score = sum in list order of w / (k + rank + 1), sorted (score total_cmp desc, MemoryId asc);
 * mem_out [n_total,2], score_out [n_total], *n_out = number of distinct ids.  n_total <= 2048. */
SG_API int sg_rrf_fuse(int device, float k, const float *weights, uint32_t n_weights, const uint64_t *mem_ids,
                       const uint32_t *list_offsets, uint32_t n_lists, uint64_t *mem_out, float *score_out,
                       uint32_t *n_out);

/* ---- sentence encoder: replaces the ONNX Runtime session + pooling inside MiniLMEmbedder
 * (src/embeddings/minilm.rs: generate_embedding_onnx :883-982, generate_embeddings_batch_onnx
 * :996-1115 = session.run + masked mean-pool + finalize_pooled :846-878), i.e. the body of
 * `impl Embedder for MiniLMEmbedder` (:1195-1377; trait src/embeddings/mod.rs:52-88).
 * Tokenisation / prefixes / metrics stay on the host, as in the reference. ------------------- */

/* weights_path: an SGW1 blob (shodh_memory_b200/encoder.py::export_bert_weights writes it from a
 * HuggingFace BERT/MiniLM state_dict: 6 layers x 384 hidden x 12 heads x 1536 FFN). */
SG_API int sg_encoder_create(sg_encoder **out, int device, const char *weights_path);
SG_API void sg_encoder_destroy(sg_encoder *e);
/* Embedder::dimension (minilm.rs:1206-1208) */
SG_API int sg_encoder_dim(const sg_encoder *e, uint32_t *dim_out);
/* Embedder::encode_batch on token ids: input_ids / attention_mask [b,L] i64 (the tensors the
 * reference binds at minilm.rs:924-934; token_type_ids are all zero there) -> emb_out [b,384]
 * f32, masked-mean-pooled and L2-normalised; all-masked rows (empty text) give zeros
 * (minilm.rs:1123-1125).  L <= 512. */
SG_API int sg_encoder_encode(sg_encoder *e, const int64_t *input_ids, const int64_t *attention_mask,
                             uint32_t b, uint32_t L, float *emb_out);
/* same with device-resident inputs/outputs, asynchronous on cuda_stream: feeds
 * sg_store_search_device without leaving the GPU (BASELINE config 5). */
SG_API int sg_encoder_encode_device(sg_encoder *e, const int64_t *d_input_ids, const int64_t *d_attention_mask,
                                    uint32_t b, uint32_t L, float *d_emb_out, void *cuda_stream);

/* The pooling tail of generate_embedding_onnx on its own (minilm.rs:959-978 masked mean-pool: pooled[d] += h[s][d]
 * for s ascending over mask == 1, / mask_sum; :846-878 finalize_pooled: NaN/Inf -> 0, norm = sqrt(sum x^2) folded
 * left to right, divide when norm > f32::EPSILON): hidden [b,L,H] f32 + attention_mask [b,L] i64 -> out [b,H],
 * bit-identical to the CPU path.  Host buffers; what sg_encoder_encode runs after its last layer. */
SG_API int sg_pool_finalize(int device, const float *hidden_f32, const int64_t *attention_mask, uint32_t b, uint32_t L,
                            uint32_t H, float *out);

#ifdef __cplusplus
}
#endif
#endif /* SHODH_GPU_H */
