//! `GpuVectorStore`: the method set RetrievalEngine calls on VamanaIndex (src/memory/retrieval.rs:53), served by
//! libshodh_gpu.so.  Same argument meaning and error behaviour as src/vector_db/vamana.rs:
//! search :764 (Ok(vec![]) on an empty index, ids dense from 0, (id, -dot) ascending, ties by id),
//! add_vector :853, build :200, mark_deleted :813, is_deleted :823, deleted_count :828, len :976.
use std::ffi::CString;
use std::path::Path;

use anyhow::{anyhow, Result};

use super::gpu_ffi::*;

pub struct GpuVectorStore {
    h: *mut SgStore,
    dim: usize,
    /// forward `search` to the graph arm (VamanaIndex::search's default) instead of the exact sweep
    pub use_graph: bool,
}

// the C library serialises calls per handle (one sweep saturates HBM; overlapping sweeps would gain nothing)
unsafe impl Send for GpuVectorStore {}
unsafe impl Sync for GpuVectorStore {}

impl GpuVectorStore {
    /// VamanaIndex::new (vamana.rs:163-198); fp16 rows on the device (768 B per 384-d vector).
    pub fn new(dim: usize, device: i32) -> Result<Self> {
        let mut h = std::ptr::null_mut();
        check(unsafe { sg_store_create(&mut h, device, dim as u32, SG_DTYPE_F16, SG_METRIC_NEG_DOT) })?;
        // reproduce the summation order of THIS build of dot_product_inline (distance_inline.rs:33)
        let order = if cfg!(target_feature = "avx2") { SG_ORDER_AVX2 } else { SG_ORDER_SCALAR };
        if let Err(e) = check(unsafe { sg_store_set_order(h, order) }) {
            unsafe { sg_store_destroy(h) };
            return Err(e);
        }
        Ok(Self { h, dim, use_graph: false })
    }

    pub fn len(&self) -> usize {
        let mut n = 0u64;
        unsafe { sg_store_len(self.h, &mut n) };
        n as usize
    }
    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }

    /// VamanaIndex::search (vamana.rs:764-808)
    pub fn search(&self, query: &[f32], k: usize) -> Result<Vec<(u32, f32)>> {
        Ok(self.search_batch(query, 1, k)?.pop().unwrap_or_default())
    }

    /// b queries at once (row-major [b, dim]); 5..=256 queries share ONE tensor-core sweep of the corpus
    pub fn search_batch(&self, queries: &[f32], b: usize, k: usize) -> Result<Vec<Vec<(u32, f32)>>> {
        if queries.len() != b * self.dim {
            return Err(anyhow!("Query dimension mismatch: expected {}, got {}", self.dim, queries.len() / b.max(1)));
        }
        if b == 0 || k == 0 {
            return Ok(vec![Vec::new(); b]);
        }
        let (mut ids, mut dist, mut cnt) = (vec![0u32; b * k], vec![0f32; b * k], vec![0u32; b]);
        let rc = unsafe {
            if self.use_graph {
                sg_store_search_ann(self.h, queries.as_ptr(), b as u32, k as u32, ids.as_mut_ptr(), dist.as_mut_ptr(),
                                    cnt.as_mut_ptr(), std::ptr::null_mut())
            } else {
                sg_store_search(self.h, queries.as_ptr(), b as u32, k as u32, ids.as_mut_ptr(), dist.as_mut_ptr(),
                                cnt.as_mut_ptr())
            }
        };
        check(rc)?;
        Ok((0..b)
            .map(|q| (0..cnt[q] as usize).map(|i| (ids[q * k + i], dist[q * k + i])).collect())
            .collect())
    }

    /// VamanaIndex::add_vector (vamana.rs:853-974) -> the new dense id
    pub fn add_vector(&mut self, vector: Vec<f32>) -> Result<u32> {
        if vector.len() != self.dim {
            return Err(anyhow!("Vector dimension mismatch: expected {}, got {}", self.dim, vector.len()));
        }
        let mut id = 0u32;
        check(unsafe { sg_store_append(self.h, vector.as_ptr(), 1, &mut id) })?;
        Ok(id)
    }

    /// VamanaIndex::build / rebuild_from_vectors (vamana.rs:200-284): one bulk upload (+ the graph when use_graph)
    pub fn build(&mut self, vectors: Vec<Vec<f32>>) -> Result<()> {
        let n = vectors.len();
        let mut flat = Vec::with_capacity(n * self.dim);
        for v in &vectors {
            if v.len() != self.dim {
                return Err(anyhow!("Vector dimension mismatch: expected {}, got {}", self.dim, v.len()));
            }
            flat.extend_from_slice(v);
        }
        check(unsafe { sg_store_reserve(self.h, (self.len() + n) as u64) })?;
        check(unsafe { sg_store_append(self.h, flat.as_ptr(), n as u64, std::ptr::null_mut()) })?;
        if self.use_graph {
            // VamanaConfig::default (vamana.rs:80-91)
            check(unsafe { sg_store_build_graph(self.h, 32, 75, 1.2, 0) })?;
        }
        Ok(())
    }

    pub fn mark_deleted(&self, id: u32) -> bool {
        let mut ok = 0;
        unsafe { sg_store_mark_deleted(self.h, id, &mut ok) == SG_OK && ok != 0 }
    }
    pub fn is_deleted(&self, id: u32) -> bool {
        let mut d = 0;
        unsafe { sg_store_is_deleted(self.h, id, &mut d) == SG_OK && d != 0 }
    }
    pub fn deleted_count(&self) -> usize {
        let mut n = 0u64;
        unsafe { sg_store_deleted_count(self.h, &mut n) };
        n as usize
    }

    pub fn extract_all_vectors(&self) -> Result<Vec<Vec<f32>>> {
        let n = self.len();
        let mut flat = vec![0f32; n * self.dim];
        check(unsafe { sg_store_get_vectors(self.h, 0, n as u64, flat.as_mut_ptr()) })?;
        Ok(flat.chunks(self.dim).map(|c| c.to_vec()).collect())
    }

    /// VamanaIndex::load_from_file (vamana_persist.rs:290-400): the reference's own VAMA v1 file, checksum verified
    pub fn load_from_file(&mut self, path: &Path) -> Result<()> {
        let c = CString::new(path.to_string_lossy().as_bytes())?;
        check(unsafe { sg_store_load_vama(self.h, c.as_ptr(), 1) })
    }
    pub fn save_to_file(&self, path: &Path) -> Result<()> {
        let c = CString::new(path.to_string_lossy().as_bytes())?;
        check(unsafe { sg_store_save_vama(self.h, c.as_ptr()) })
    }

    /// IdMapping (retrieval.rs:64-140) for search_ids: vec2mem[i] = Uuid of the memory owning vector i
    pub fn set_memory_ids(&self, vec2mem: &[uuid::Uuid]) -> Result<()> {
        let mut flat = Vec::with_capacity(vec2mem.len() * 2);
        for u in vec2mem {
            let (hi, lo) = u.as_u64_pair();
            flat.push(hi);
            flat.push(lo);
        }
        check(unsafe { sg_store_set_memory_ids(self.h, flat.as_ptr(), vec2mem.len() as u64) })
    }

    /// RetrievalEngine::search_ids after the embed step (retrieval.rs:905-964): 4*limit candidates from the scan,
    /// similarity = -distance, best chunk per MemoryId, (similarity desc, MemoryId asc), truncate -- all on the device
    pub fn search_ids(&self, query: &[f32], limit: usize) -> Result<Vec<(uuid::Uuid, f32)>> {
        if query.len() != self.dim {
            return Err(anyhow!("Query dimension mismatch: expected {}, got {}", self.dim, query.len()));
        }
        let (mut mem, mut sim, mut cnt) = (vec![0u64; limit * 2], vec![0f32; limit], 0u32);
        check(unsafe {
            sg_store_search_memory_ids(self.h, query.as_ptr(), 1, limit as u32, mem.as_mut_ptr(), sim.as_mut_ptr(), &mut cnt)
        })?;
        Ok((0..cnt as usize).map(|i| (uuid::Uuid::from_u64_pair(mem[2 * i], mem[2 * i + 1]), sim[i])).collect())
    }
}

impl Drop for GpuVectorStore {
    fn drop(&mut self) {
        unsafe { sg_store_destroy(self.h) }
    }
}

/// One candidate of RelevanceEngine's ranking step (the seven signals fuse_scores_full reads, relevance.rs:529-594)
pub struct FuseCandidate {
    pub semantic: f32,
    pub entity: f32,
    pub tag: f32,
    pub importance: f32,
    pub momentum: f32,
    pub access_count: u32,
    pub graph_strength: f32,
    pub created_at_unix_s: i64,
}

/// src/relevance.rs:805-918 -- fuse_scores_full + calibrate_score + recency boost + graph boost + sort
/// (score desc, created_at desc, position asc) + `>= min_score` + truncate, as ONE device call.
/// Returns (candidate position, fused score) in final order.
pub fn fuse_and_rank(device: i32, w: &SgFuseWeights, cfg: &SgFuseConfig, c: &[FuseCandidate], now_unix_s: i64)
                     -> Result<Vec<(usize, f32)>> {
    let n = c.len();
    if n == 0 || cfg.max_results == 0 {
        return Ok(Vec::new());
    }
    let col = |f: fn(&FuseCandidate) -> f32| c.iter().map(f).collect::<Vec<f32>>();
    let (sem, ent, tag) = (col(|x| x.semantic), col(|x| x.entity), col(|x| x.tag));
    let (imp, mom, gr) = (col(|x| x.importance), col(|x| x.momentum), col(|x| x.graph_strength));
    let acc: Vec<u32> = c.iter().map(|x| x.access_count).collect();
    let created: Vec<i64> = c.iter().map(|x| x.created_at_unix_s).collect();
    let feats = SgCandidateFeats {
        semantic: sem.as_ptr(), entity: ent.as_ptr(), tag: tag.as_ptr(), importance: imp.as_ptr(), momentum: mom.as_ptr(),
        access_count: acc.as_ptr(), graph_strength: gr.as_ptr(), created_at_unix_s: created.as_ptr(),
    };
    let m = cfg.max_results as usize;
    let (mut order, mut score, mut n_out) = (vec![0u32; m], vec![0f32; m], 0u32);
    check(unsafe {
        sg_fuse_topk(device, w, cfg, &feats, n as u32, now_unix_s, order.as_mut_ptr(), score.as_mut_ptr(), &mut n_out,
                     std::ptr::null_mut())
    })?;
    Ok((0..n_out as usize).map(|i| (order[i] as usize, score[i])).collect())
}
