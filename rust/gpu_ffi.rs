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
