//! `GpuCompressedStore` / `GpuSpannIndex`: the method sets of `CompressedVectorStore` (src/vector_db/pq.rs:400-466) and
//! `SpannIndex` (src/vector_db/spann.rs:363-1051) over libshodh_gpu.so.  Feature `gpu-b200`; goes to `src/vector_db/gpu_pq.rs`.
//! Distances and ids equal the CPU implementation's bit for bit (same f32 operations in the same order), so either side
//! can serve a query.  NOT COMPILED in this repo (no Rust toolchain in the build image) -- see rust/README.md.
use super::gpu_ffi::*;
use anyhow::Result;
use std::ffi::CString;
use std::ptr;

pub struct GpuCompressedStore {
    h: *mut SgPq,
}
unsafe impl Send for GpuCompressedStore {}
unsafe impl Sync for GpuCompressedStore {} // the library serialises calls on one handle

impl GpuCompressedStore {
    /// `ProductQuantizer::train` (pq.rs:94-215).  `orders[m]` is the shuffled index order `kmeans` draws for subspace m
    /// (`indices.shuffle(&mut thread_rng())`, :157-159): pass the same orders to get the CPU's centroids exactly.
    pub fn train(device: i32, dim: usize, subvec_dim: usize, num_centroids: usize, iterations: usize, vectors: &[Vec<f32>],
                 orders: &[Vec<u32>]) -> Result<Self> {
        let n = vectors.len();
        let flat: Vec<f32> = vectors.iter().flat_map(|v| v.iter().copied()).collect();
        let ord: Vec<u32> = orders.iter().flat_map(|o| o.iter().copied()).collect();
        anyhow::ensure!(flat.len() == n * dim && ord.len() == (dim / subvec_dim) * n, "shape mismatch");
        let mut h = ptr::null_mut();
        check(unsafe {
            sg_pq_train(&mut h, device, dim as u32, subvec_dim as u32, num_centroids as u32, iterations as u32, flat.as_ptr(),
                        n as u64, ord.as_ptr())
        })?;
        Ok(Self { h })
    }

    /// from trained centroids `[dim/subvec_dim][num_centroids][subvec_dim]` (e.g. a quantizer loaded from disk)
    pub fn from_centroids(device: i32, dim: usize, subvec_dim: usize, num_centroids: usize, centroids: &[f32]) -> Result<Self> {
        let mut h = ptr::null_mut();
        check(unsafe { sg_pq_create(&mut h, device, dim as u32, subvec_dim as u32, num_centroids as u32, centroids.as_ptr()) })?;
        Ok(Self { h })
    }

    /// `CompressedVectorStore::add` (pq.rs:428-432) for a batch
    pub fn add(&self, ids: &[u32], vectors: &[f32]) -> Result<()> {
        check(unsafe { sg_pq_add(self.h, ids.as_ptr(), vectors.as_ptr(), ids.len() as u64) })
    }

    /// `ProductQuantizer::encode` (pq.rs:220-257)
    pub fn encode(&self, vectors: &[f32], n: usize, num_subvectors: usize) -> Result<Vec<u8>> {
        let mut codes = vec![0u8; n * num_subvectors];
        check(unsafe { sg_pq_encode(self.h, vectors.as_ptr(), n as u64, codes.as_mut_ptr()) })?;
        Ok(codes)
    }

    /// `CompressedVectorStore::search` (pq.rs:449-466) for `b` queries at once: `(id, distance)` ascending by
    /// `(distance, id)`, exactly the CPU result
    pub fn search_batch(&self, queries: &[f32], b: usize, k: usize) -> Result<Vec<Vec<(u32, f32)>>> {
        let (mut ids, mut dist, mut cnt) = (vec![0u32; b * k], vec![0f32; b * k], vec![0u32; b]);
        check(unsafe { sg_pq_search(self.h, queries.as_ptr(), b as u32, k as u32, ids.as_mut_ptr(), dist.as_mut_ptr(), cnt.as_mut_ptr()) })?;
        Ok((0..b).map(|q| (0..cnt[q] as usize).map(|i| (ids[q * k + i], dist[q * k + i])).collect()).collect())
    }
}
impl Drop for GpuCompressedStore {
    fn drop(&mut self) {
        unsafe { sg_pq_destroy(self.h) }
    }
}

pub struct GpuSpannIndex {
    h: *mut SgSpann,
}
unsafe impl Send for GpuSpannIndex {}
unsafe impl Sync for GpuSpannIndex {}

impl GpuSpannIndex {
    /// `SpannIndex::load_from_file` (spann.rs:891-1003): the reference's own `SPAN` v1 file
    pub fn load_from_file(device: i32, path: &std::path::Path, num_probes: usize, verify_checksum: bool) -> Result<Self> {
        let c = CString::new(path.to_string_lossy().as_bytes())?;
        let mut h = ptr::null_mut();
        check(unsafe { sg_spann_load_file(&mut h, device, c.as_ptr(), num_probes as u32, verify_checksum as i32) })?;
        Ok(Self { h })
    }
    /// `SpannIndex::save_to_file` (spann.rs:750-888): byte-identical layout, loadable by the reference
    pub fn save_to_file(&self, path: &std::path::Path) -> Result<()> {
        let c = CString::new(path.to_string_lossy().as_bytes())?;
        check(unsafe { sg_spann_save_file(self.h, c.as_ptr()) })
    }
    /// `SpannIndex::insert` (spann.rs:1006-1051)
    pub fn insert(&self, vector_id: u32, vector: &[f32]) -> Result<()> {
        check(unsafe { sg_spann_insert(self.h, vector_id, vector.as_ptr()) })
    }
    /// `SpannIndex::search` (spann.rs:574-693) for `b` queries at once
    pub fn search_batch(&self, queries: &[f32], b: usize, k: usize) -> Result<Vec<Vec<(u32, f32)>>> {
        let (mut ids, mut dist, mut cnt) = (vec![0u32; b * k], vec![0f32; b * k], vec![0u32; b]);
        check(unsafe { sg_spann_search(self.h, queries.as_ptr(), b as u32, k as u32, ids.as_mut_ptr(), dist.as_mut_ptr(), cnt.as_mut_ptr()) })?;
        Ok((0..b).map(|q| (0..cnt[q] as usize).map(|i| (ids[q * k + i], dist[q * k + i])).collect()).collect())
    }
}
impl Drop for GpuSpannIndex {
    fn drop(&mut self) {
        unsafe { sg_spann_destroy(self.h) }
    }
}
