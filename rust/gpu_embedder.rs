//! `impl Embedder for GpuMiniLM` (src/embeddings/mod.rs:52-88): the tokeniser stays on the host exactly as in
//! MiniLMEmbedder (src/embeddings/minilm.rs:901-921); `session.run` + masked mean-pool + finalize_pooled
//! (:939-978, :846-878) are replaced by ONE call, sg_encoder_encode.  No hash fallback: if the device or the
//! weights are missing construction fails (contrast minilm.rs:1182-1190).
use std::ffi::CString;
use std::path::Path;

use anyhow::{anyhow, Result};
use tokenizers::Tokenizer;

use super::Embedder;
use crate::vector_db::gpu_ffi::*;

pub struct GpuMiniLM {
    h: *mut SgEncoder,
    tokenizer: Tokenizer,
    dim: usize,
    max_len: usize,
}

unsafe impl Send for GpuMiniLM {}
unsafe impl Sync for GpuMiniLM {}

impl GpuMiniLM {
    /// `weights`: the SGW1 blob written once by shodh_memory_b200.encoder.export_bert_weights from the
    /// all-MiniLM-L6-v2 checkpoint; `tokenizer_json`: the file MiniLMEmbedder already loads (minilm.rs:430-470).
    pub fn new(weights: &Path, tokenizer_json: &Path, device: i32) -> Result<Self> {
        let tokenizer = Tokenizer::from_file(tokenizer_json).map_err(|e| anyhow!("tokenizer: {e}"))?;
        let c = CString::new(weights.to_string_lossy().as_bytes())?;
        let mut h = std::ptr::null_mut();
        check(unsafe { sg_encoder_create(&mut h, device, c.as_ptr()) })?;
        let mut dim = 0u32;
        check(unsafe { sg_encoder_dim(h, &mut dim) })?;
        Ok(Self { h, tokenizer, dim: dim as usize, max_len: 256 })   // max_length 256: minilm.rs:225
    }

    fn encode_texts(&self, texts: &[&str]) -> Result<Vec<Vec<f32>>> {
        if texts.is_empty() {
            return Ok(Vec::new());
        }
        let enc = self.tokenizer.encode_batch(texts.to_vec(), true).map_err(|e| anyhow!("tokenize: {e}"))?;
        // pad only to the longest sequence of the batch: the fp16 graph is pad-invariant
        let l = enc.iter().map(|e| e.get_ids().len().min(self.max_len)).max().unwrap_or(1).max(1);
        let b = texts.len();
        let (mut ids, mut mask) = (vec![0i64; b * l], vec![0i64; b * l]);
        for (i, e) in enc.iter().enumerate() {
            // empty text -> all-masked row -> 384 zeros (minilm.rs:1123-1125)
            if texts[i].trim().is_empty() {
                continue;
            }
            for (j, (&t, &m)) in e.get_ids().iter().zip(e.get_attention_mask()).take(l).enumerate() {
                ids[i * l + j] = t as i64;
                mask[i * l + j] = m as i64;
            }
        }
        let mut out = vec![0f32; b * self.dim];
        check(unsafe { sg_encoder_encode(self.h, ids.as_ptr(), mask.as_ptr(), b as u32, l as u32, out.as_mut_ptr()) })?;
        Ok(out.chunks(self.dim).map(|c| c.to_vec()).collect())
    }
}

impl Embedder for GpuMiniLM {
    fn encode(&self, text: &str) -> Result<Vec<f32>> {
        Ok(self.encode_texts(&[text])?.pop().unwrap_or_else(|| vec![0.0; self.dim]))
    }
    fn encode_batch(&self, texts: &[&str]) -> Result<Vec<Vec<f32>>> {
        self.encode_texts(texts)
    }
    fn dimension(&self) -> usize {
        self.dim
    }
    /// full encoded length incl. special tokens, without truncation (what MiniLMEmbedder overrides too)
    fn count_tokens(&self, text: &str) -> usize {
        match self.tokenizer.encode(text, true) {
            Ok(e) => e.get_ids().len(),
            Err(_) => crate::token_estimation::estimate_tokens(text) + super::chunking::SPECIAL_TOKEN_OVERHEAD,
        }
    }
}

impl Drop for GpuMiniLM {
    fn drop(&mut self) {
        unsafe { sg_encoder_destroy(self.h) }
    }
}
