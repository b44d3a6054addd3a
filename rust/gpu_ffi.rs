//! FFI declarations for libshodh_gpu.so (include/shodh_gpu.h).  Feature `gpu-b200`.
//! Build: `cargo:rustc-link-lib=dylib=shodh_gpu` + `cargo:rustc-link-search=native=$SHODH_GPU_LIB_DIR`.
#![allow(non_camel_case_types, dead_code)]

use std::os::raw::{c_char, c_int, c_void};

pub const SG_OK: c_int = 0;
pub const SG_ERR_INVALID: c_int = -1;
pub const SG_ERR_CUDA: c_int = -2;
pub const SG_ERR_OOM: c_int = -3;
pub const SG_ERR_DIM: c_int = -4;
pub const SG_ERR_UNSUPPORTED: c_int = -5;
pub const SG_ERR_NO_DEVICE: c_int = -6;
pub const SG_ERR_NOT_BUILT: c_int = -7;

pub const SG_DTYPE_F32: c_int = 0;
pub const SG_DTYPE_F16: c_int = 1;
/// `DistanceMetric as c_int` (src/vector_db/vamana.rs:38-54)
pub const SG_METRIC_NEG_DOT: c_int = 0;
pub const SG_METRIC_EUCLIDEAN: c_int = 1;
pub const SG_METRIC_COSINE: c_int = 2;
/// summation order of THIS build of distance_inline.rs (`cfg(target_feature = "avx2")` at :33)
pub const SG_ORDER_SCALAR: c_int = 0;
pub const SG_ORDER_AVX2: c_int = 1;

#[repr(C)]
pub struct SgStore {
    _private: [u8; 0],
}
#[repr(C)]
pub struct SgEncoder {
    _private: [u8; 0],
}
#[repr(C)]
pub struct SgPq {
    _private: [u8; 0],
}
#[repr(C)]
pub struct SgSpann {
    _private: [u8; 0],
}

/// LearnedWeights (src/relevance.rs:64-88)
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct SgFuseWeights {
    pub semantic: f32,
    pub entity: f32,
    pub tag: f32,
    pub importance: f32,
    pub momentum: f32,
    pub access_count: f32,
    pub graph_strength: f32,
}

/// the RelevanceConfig fields the ranking reads (src/relevance.rs:147-177, :914)
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct SgFuseConfig {
    pub recency_boost_hours: u64,
    pub recency_boost_multiplier: f32,
    pub graph_boost_multiplier: f32,
    pub min_score: f32,
    pub max_results: u32,
}

/// structure-of-arrays view of the candidate memories (one entry per candidate in every array)
#[repr(C)]
pub struct SgCandidateFeats {
    pub semantic: *const f32,
    pub entity: *const f32,
    pub tag: *const f32,
    pub importance: *const f32,
    pub momentum: *const f32,
    pub access_count: *const u32,
    pub graph_strength: *const f32,
    pub created_at_unix_s: *const i64,
}

extern "C" {
    pub fn sg_version() -> *const c_char;
    pub fn sg_last_error() -> *const c_char;
    pub fn sg_device_count(count_out: *mut c_int) -> c_int;

    // ---- vector store: VamanaIndex as RetrievalEngine holds it (retrieval.rs:53) ----
    pub fn sg_store_create(out: *mut *mut SgStore, device: c_int, dim: u32, dtype: c_int, metric: c_int) -> c_int;
    pub fn sg_store_destroy(s: *mut SgStore);
    pub fn sg_store_reserve(s: *mut SgStore, n_rows: u64) -> c_int;
    pub fn sg_store_append(s: *mut SgStore, rows: *const f32, n: u64, first_id: *mut u32) -> c_int;
    pub fn sg_store_append_device(s: *mut SgStore, d_rows: *const f32, n: u64, first_id: *mut u32, stream: *mut c_void) -> c_int;
    pub fn sg_store_mark_deleted(s: *mut SgStore, id: u32, ok: *mut c_int) -> c_int;
    pub fn sg_store_is_deleted(s: *mut SgStore, id: u32, deleted: *mut c_int) -> c_int;
    pub fn sg_store_clear_deleted(s: *mut SgStore) -> c_int;
    pub fn sg_store_len(s: *const SgStore, n: *mut u64) -> c_int;
    pub fn sg_store_deleted_count(s: *const SgStore, n: *mut u64) -> c_int;
    pub fn sg_store_get_vectors(s: *mut SgStore, first: u64, n: u64, out: *mut f32) -> c_int;
    pub fn sg_store_set_order(s: *mut SgStore, order: c_int) -> c_int;
    pub fn sg_store_load_vama(s: *mut SgStore, path: *const c_char, verify_checksum: c_int) -> c_int;
    pub fn sg_store_save_vama(s: *mut SgStore, path: *const c_char) -> c_int;
    /// VamanaIndex::search on its exact arm (brute_force_search, vamana.rs:1167-1188), b queries at once
    pub fn sg_store_search(s: *mut SgStore, queries: *const f32, b: u32, k: u32, ids: *mut u32, dist: *mut f32,
                           counts: *mut u32) -> c_int;
    /// RetrievalEngine::search_ids from the embedding on (retrieval.rs:905-964), de-dup + re-rank on the device
    pub fn sg_store_set_memory_ids(s: *mut SgStore, vec2mem: *const u64, n: u64) -> c_int;
    pub fn sg_store_search_memory_ids(s: *mut SgStore, queries: *const f32, b: u32, limit: u32, mem_out: *mut u64,
                                      sim_out: *mut f32, counts: *mut u32) -> c_int;
    // graph arm (VamanaIndex::search without SHODH_VECTOR_EXACT)
    pub fn sg_store_build_graph(s: *mut SgStore, max_degree: u32, search_list_size: u32, alpha: f32, seed: u64) -> c_int;
    pub fn sg_store_enable_graph(s: *mut SgStore, max_degree: u32, search_list_size: u32, alpha: f32) -> c_int;
    pub fn sg_store_search_ann(s: *mut SgStore, queries: *const f32, b: u32, k: u32, ids: *mut u32, dist: *mut f32,
                               counts: *mut u32, stats: *mut u64) -> c_int;

    // ---- similarity.rs ----
    pub fn sg_top_k_similar(device: c_int, query: *const f32, query_dim: u32, candidates: *const f32, n: u64, dim: u32,
                            k: u32, order: c_int, pos: *mut u32, score: *mut f32, n_out: *mut u32) -> c_int;

    // ---- relevance.rs / hybrid_search.rs ----
    pub fn sg_fuse_default_weights(w: *mut SgFuseWeights);
    pub fn sg_fuse_default_config(c: *mut SgFuseConfig);
    pub fn sg_fuse_topk(device: c_int, w: *const SgFuseWeights, cfg: *const SgFuseConfig, feats: *const SgCandidateFeats,
                        n: u32, now_unix_s: i64, order: *mut u32, score: *mut f32, n_out: *mut u32, all: *mut f32) -> c_int;
    pub fn sg_rrf_fuse(device: c_int, k: f32, weights: *const f32, n_weights: u32, mem_ids: *const u64,
                       list_offsets: *const u32, n_lists: u32, mem_out: *mut u64, score_out: *mut f32, n_out: *mut u32) -> c_int;

    // ---- embeddings: Embedder (src/embeddings/mod.rs:52-88) ----
    pub fn sg_encoder_create(out: *mut *mut SgEncoder, device: c_int, weights_path: *const c_char) -> c_int;
    pub fn sg_encoder_destroy(e: *mut SgEncoder);
    pub fn sg_encoder_dim(e: *const SgEncoder, dim_out: *mut u32) -> c_int;
    pub fn sg_encoder_encode(e: *mut SgEncoder, input_ids: *const i64, attention_mask: *const i64, b: u32, l: u32,
                             emb_out: *mut f32) -> c_int;
    pub fn sg_pool_finalize(device: c_int, hidden: *const f32, attention_mask: *const i64, b: u32, l: u32, h: u32,
                            out: *mut f32) -> c_int;

    // ---- store diagnostics / knobs ----
    pub fn sg_store_dim(s: *const SgStore, dim_out: *mut u32) -> c_int;
    pub fn sg_store_metric(s: *const SgStore, metric_out: *mut c_int) -> c_int;
    /// in_place = 1: rows live in a reserved address range that grows by mapping chunks (no copy of the corpus)
    pub fn sg_store_memory_info(s: *mut SgStore, in_place: *mut c_int, capacity_rows: *mut u64, bytes: *mut u64,
                                chunks: *mut u32) -> c_int;
    /// queries served by the tensor-core sweep / of those, re-run on the exact sweep (results are exact either way)
    pub fn sg_store_tc_stats(s: *mut SgStore, queries: *mut u64, flagged: *mut u64) -> c_int;
    /// 0 = automatic (tensor-core sweep for batches), 1 = exact sweep only
    pub fn sg_store_set_tc_mode(s: *mut SgStore, mode: c_int) -> c_int;
    pub fn sg_store_set_id_base(s: *mut SgStore, id_base: u32) -> c_int;
    pub fn sg_store_set_graph(s: *mut SgStore, max_degree: u32, medoid: u32, deg: *const u32, nbr: *const u32,
                              stride: u32, n_nodes: u64) -> c_int;
    pub fn sg_store_get_graph(s: *mut SgStore, deg_out: *mut u32, nbr_out: *mut u32, n_nodes: u64) -> c_int;
    pub fn sg_store_graph_info(s: *mut SgStore, max_degree: *mut u32, medoid: *mut u32, graph_len: *mut u64) -> c_int;

    // ---- pq.rs: ProductQuantizer + CompressedVectorStore (src/vector_db/pq.rs:94-466) ----
    pub fn sg_pq_create(out: *mut *mut SgPq, device: c_int, dim: u32, subvec_dim: u32, num_centroids: u32,
                        centroids: *const f32) -> c_int;
    /// `orders` [dim/subvec_dim][n]: the shuffled index orders `kmeans` draws from `thread_rng` (pq.rs:157-159)
    pub fn sg_pq_train(out: *mut *mut SgPq, device: c_int, dim: u32, subvec_dim: u32, num_centroids: u32, iterations: u32,
                       vectors: *const f32, n: u64, orders: *const u32) -> c_int;
    pub fn sg_pq_destroy(p: *mut SgPq);
    pub fn sg_pq_info(p: *mut SgPq, dim: *mut u32, num_subvectors: *mut u32, num_centroids: *mut u32, len: *mut u64) -> c_int;
    pub fn sg_pq_get_centroids(p: *mut SgPq, centroids_out: *mut f32) -> c_int;
    pub fn sg_pq_encode(p: *mut SgPq, rows: *const f32, n: u64, codes_out: *mut u8) -> c_int;
    pub fn sg_pq_decode(p: *mut SgPq, codes: *const u8, n: u64, rows_out: *mut f32) -> c_int;
    pub fn sg_pq_table(p: *mut SgPq, query: *const f32, table_out: *mut f32) -> c_int;
    pub fn sg_pq_add(p: *mut SgPq, ids: *const u32, rows: *const f32, n: u64) -> c_int;
    pub fn sg_pq_search(p: *mut SgPq, queries: *const f32, b: u32, k: u32, ids: *mut u32, dist: *mut f32,
                        counts: *mut u32) -> c_int;

    // ---- spann.rs: SpannIndex (src/vector_db/spann.rs:363-1083) ----
    pub fn sg_spann_build(out: *mut *mut SgSpann, device: c_int, dim: u32, num_partitions: u32, num_probes: u32,
                          kmeans_iterations: u32, metric: c_int, vectors: *const f32, n: u64, order: *const u32,
                          pq_iterations: u32, pq_orders: *const u32) -> c_int;
    pub fn sg_spann_destroy(s: *mut SgSpann);
    pub fn sg_spann_info(s: *mut SgSpann, num_partitions: *mut u32, len: *mut u64, num_probes: *mut u32) -> c_int;
    pub fn sg_spann_set_probes(s: *mut SgSpann, num_probes: u32) -> c_int;
    pub fn sg_spann_load_file(out: *mut *mut SgSpann, device: c_int, path: *const c_char, num_probes: u32,
                              verify_checksum: c_int) -> c_int;
    pub fn sg_spann_save_file(s: *mut SgSpann, path: *const c_char) -> c_int;
    pub fn sg_spann_insert(s: *mut SgSpann, vector_id: u32, vector: *const f32) -> c_int;
    pub fn sg_spann_search(s: *mut SgSpann, queries: *const f32, b: u32, k: u32, ids: *mut u32, dist: *mut f32,
                           counts: *mut u32) -> c_int;
}

/// negative status -> anyhow error carrying the library's message (never panics, never aborts)
pub fn check(rc: c_int) -> anyhow::Result<()> {
    if rc == SG_OK {
        return Ok(());
    }
    let msg = unsafe {
        let p = sg_last_error();
        if p.is_null() { String::new() } else { std::ffi::CStr::from_ptr(p).to_string_lossy().into_owned() }
    };
    Err(anyhow::anyhow!("libshodh_gpu error {rc}: {msg}"))
}
