// shodh_gpu.hpp -- C++ host-side mirror of the reference's Rust surface over the C ABI.
//
// The reference is compiled Rust and this image has no Rust toolchain, so the host layer above
// include/shodh_gpu.h is C++ (header-only), with the reference's names, argument meaning and
// error behaviour:
//   shodh::GpuVectorStore  <->  VamanaIndex as RetrievalEngine uses it (src/vector_db/vamana.rs:
//                               search :764, add_vector :853, build :200, mark_deleted :813,
//                               is_deleted :823, deleted_count :828, len :976, extract_all_vectors)
//   shodh::GpuMiniLM       <->  impl Embedder for MiniLMEmbedder on token ids
//                               (src/embeddings/mod.rs:52-88, src/embeddings/minilm.rs:1195-1377)
//   shodh::fuse_and_rank   <->  ranking phase of RelevanceEngine::surface_relevant_inner
//                               (src/relevance.rs:805-918)
// Errors surface as shodh::Error (the analogue of anyhow::Result); nothing here computes on the CPU.
// INTEGRATION.md shows the equivalent Rust `extern "C"` wrapper.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/shodh_gpu.h"

namespace shodh {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc)
{
    if (rc != SG_OK) throw Error(rc, std::string("libshodh_gpu: ") + sg_last_error());
}

class GpuVectorStore {
public:
    /// metric = VamanaConfig::distance_metric (vamana.rs:38-54); RetrievalEngine uses NormalizedDotProduct
    explicit GpuVectorStore(uint32_t dim = 384, sg_dtype dtype = SG_DTYPE_F16, int device = 0,
                            sg_metric metric = SG_METRIC_NEG_DOT) : dim_(dim)
    {
        check(sg_store_create(&h_, device, dim, dtype, metric));
    }
    ~GpuVectorStore() { sg_store_destroy(h_); }
    GpuVectorStore(const GpuVectorStore &) = delete;
    GpuVectorStore &operator=(const GpuVectorStore &) = delete;

    /// VamanaIndex::len
    size_t len() const { uint64_t n = 0; check(sg_store_len(h_, &n)); return (size_t)n; }
    /// how the row matrix is held: grows in place (CUDA virtual-memory chunks) or by allocate-and-copy; rows that fit
    struct MemoryInfo { bool in_place; uint64_t capacity_rows, bytes; uint32_t chunks; };
    MemoryInfo memory_info() const
    {
        int ip = 0; uint64_t cap = 0, by = 0; uint32_t ch = 0;
        check(sg_store_memory_info(h_, &ip, &cap, &by, &ch));
        return MemoryInfo{ip != 0, cap, by, ch};
    }
    bool is_empty() const { return len() == 0; }
    /// VamanaIndex::add_vector -> the new (dense, sequential) id
    uint32_t add_vector(const std::vector<float> &v)
    {
        if (v.size() != dim_) throw Error(SG_ERR_DIM, "dimension mismatch");
        uint32_t id = 0;
        check(sg_store_append(h_, v.data(), 1, &id));
        return id;
    }
    /// VamanaIndex::build / rebuild_from_vectors: one bulk upload of row-major [n,dim]
    void build(const std::vector<float> &rows_flat)
    {
        if (rows_flat.size() % dim_ != 0) throw Error(SG_ERR_DIM, "dimension mismatch");
        check(sg_store_append(h_, rows_flat.data(), rows_flat.size() / dim_, nullptr));
    }
    bool mark_deleted(uint32_t id) { int ok = 0; check(sg_store_mark_deleted(h_, id, &ok)); return ok != 0; }
    bool is_deleted(uint32_t id) { int d = 0; check(sg_store_is_deleted(h_, id, &d)); return d != 0; }
    size_t deleted_count() const { uint64_t n = 0; check(sg_store_deleted_count(h_, &n)); return (size_t)n; }
    void clear_deleted() { check(sg_store_clear_deleted(h_)); }
    std::vector<float> extract_all_vectors()
    {
        std::vector<float> out(len() * dim_);
        if (!out.empty()) check(sg_store_get_vectors(h_, 0, len(), out.data()));
        return out;
    }
    /// which build of distance_inline.rs to reproduce (scalar order by default)
    void set_order(sg_order o) { check(sg_store_set_order(h_, o)); }

    /// VamanaIndex::search: (id, distance = -dot) ascending, ties by id; empty store -> empty vector
    std::vector<std::pair<uint32_t, float>> search(const std::vector<float> &query, size_t k)
    {
        if (query.size() != dim_) throw Error(SG_ERR_DIM, "dimension mismatch");
        auto r = search_batch(query, 1, k);
        return r.empty() ? std::vector<std::pair<uint32_t, float>>{} : r[0];
    }
    /// b queries (row-major [b,dim]) in one call: 5..256 queries share one corpus sweep
    std::vector<std::vector<std::pair<uint32_t, float>>> search_batch(const std::vector<float> &queries, size_t b, size_t k)
    {
        if (queries.size() != b * dim_) throw Error(SG_ERR_DIM, "dimension mismatch");
        std::vector<uint32_t> ids(b * k), cnt(b);
        std::vector<float> dist(b * k);
        check(sg_store_search(h_, queries.data(), (uint32_t)b, (uint32_t)k, ids.data(), dist.data(), cnt.data()));
        std::vector<std::vector<std::pair<uint32_t, float>>> out(b);
        for (size_t q = 0; q < b; ++q)
            for (uint32_t j = 0; j < cnt[q]; ++j) out[q].emplace_back(ids[q * k + j], dist[q * k + j]);
        return out;
    }
    /// VamanaIndex::build's graph half over the rows already uploaded (VamanaConfig::default: 32, 75, 1.2)
    void build_graph(uint32_t max_degree = 32, uint32_t search_list_size = 75, float alpha = 1.2f, uint64_t seed = 0)
    {
        check(sg_store_build_graph(h_, max_degree, search_list_size, alpha, seed));
    }
    /// an empty index that links every add_vector like the reference does (vamana.rs:853-973)
    void enable_graph(uint32_t max_degree = 32, uint32_t search_list_size = 75, float alpha = 1.2f)
    {
        check(sg_store_enable_graph(h_, max_degree, search_list_size, alpha));
    }
    /// VamanaIndex::search without SHODH_VECTOR_EXACT: beam search over the graph (vamana.rs:779-807);
    /// throws Error(SG_ERR_NOT_BUILT) where the reference says "Vamana graph not built"
    std::vector<std::pair<uint32_t, float>> search_ann(const std::vector<float> &query, size_t k)
    {
        if (query.size() != dim_) throw Error(SG_ERR_DIM, "dimension mismatch");
        std::vector<uint32_t> ids(k);
        std::vector<float> dist(k);
        uint32_t cnt = 0;
        check(sg_store_search_ann(h_, query.data(), 1, (uint32_t)k, ids.data(), dist.data(), &cnt, nullptr));
        std::vector<std::pair<uint32_t, float>> out;
        for (uint32_t j = 0; j < cnt; ++j) out.emplace_back(ids[j], dist[j]);
        return out;
    }
    sg_store *handle() { return h_; }

private:
    sg_store *h_ = nullptr;
    uint32_t dim_;
};

/// ProductQuantizer + CompressedVectorStore (src/vector_db/pq.rs) over sg_pq_*: codes, tables and the
/// (distance, id) ranking are bit-identical to the CPU path
class GpuProductQuantizer {
public:
    /// from trained centroids, flattened [num_subvectors][num_centroids][subvec_dim]
    GpuProductQuantizer(const std::vector<float> &centroids, uint32_t dim, uint32_t subvec_dim, uint32_t num_centroids, int device = 0)
        : dim_(dim), m_(dim / subvec_dim)
    {
        if (centroids.size() != (size_t)dim * num_centroids) throw Error(SG_ERR_DIM, "centroid array has the wrong size");
        check(sg_pq_create(&h_, device, dim, subvec_dim, num_centroids, centroids.data()));
    }
    ~GpuProductQuantizer() { sg_pq_destroy(h_); }
    GpuProductQuantizer(const GpuProductQuantizer &) = delete;
    GpuProductQuantizer &operator=(const GpuProductQuantizer &) = delete;
    /// ProductQuantizer::encode
    std::vector<uint8_t> encode(const std::vector<float> &v)
    {
        if (v.size() != dim_) throw Error(SG_ERR_DIM, "dimension mismatch");
        std::vector<uint8_t> codes(m_);
        check(sg_pq_encode(h_, v.data(), 1, codes.data()));
        return codes;
    }
    /// CompressedVectorStore::add
    void add(uint32_t id, const std::vector<float> &v)
    {
        if (v.size() != dim_) throw Error(SG_ERR_DIM, "dimension mismatch");
        check(sg_pq_add(h_, &id, v.data(), 1));
    }
    /// CompressedVectorStore::search -> (id, distance) ascending by (distance, id)
    std::vector<std::pair<uint32_t, float>> search(const std::vector<float> &query, size_t k)
    {
        if (query.size() != dim_) throw Error(SG_ERR_DIM, "dimension mismatch");
        std::vector<uint32_t> ids(k);
        std::vector<float> dist(k);
        uint32_t cnt = 0;
        check(sg_pq_search(h_, query.data(), 1, (uint32_t)k, ids.data(), dist.data(), &cnt));
        std::vector<std::pair<uint32_t, float>> out;
        for (uint32_t j = 0; j < cnt; ++j) out.emplace_back(ids[j], dist[j]);
        return out;
    }
    size_t len() { uint64_t n = 0; check(sg_pq_info(h_, nullptr, nullptr, nullptr, &n)); return (size_t)n; }

private:
    sg_pq *h_ = nullptr;
    uint32_t dim_, m_;
};

/// similarity.rs::top_k_similar over a dense candidate matrix: (score, position) by score descending, ties in
/// input order; cosine_similarity is 0 on a zero norm or a length mismatch (similarity.rs:10-48)
inline std::vector<std::pair<float, uint32_t>> top_k_similar(const std::vector<float> &query, const std::vector<float> &candidates_flat,
                                                              size_t dim, size_t k, int device = 0, sg_order order = SG_ORDER_SCALAR)
{
    const size_t n = dim ? candidates_flat.size() / dim : 0;
    std::vector<uint32_t> pos(k ? k : 1);
    std::vector<float> sc(k ? k : 1);
    uint32_t r = 0;
    check(sg_top_k_similar(device, query.data(), (uint32_t)query.size(), candidates_flat.data(), n, (uint32_t)dim, (uint32_t)k,
                           order, pos.data(), sc.data(), &r));
    std::vector<std::pair<float, uint32_t>> out;
    for (uint32_t j = 0; j < r; ++j) out.emplace_back(sc[j], pos[j]);
    return out;
}

class GpuMiniLM {
public:
    explicit GpuMiniLM(const std::string &weights_path, int device = 0) { check(sg_encoder_create(&h_, device, weights_path.c_str())); }
    ~GpuMiniLM() { sg_encoder_destroy(h_); }
    GpuMiniLM(const GpuMiniLM &) = delete;
    GpuMiniLM &operator=(const GpuMiniLM &) = delete;
    /// Embedder::dimension
    size_t dimension() const { uint32_t d = 0; check(sg_encoder_dim(h_, &d)); return d; }
    /// Embedder::encode_batch after tokenisation: ids/mask are [b,L] row-major
    std::vector<float> encode_ids(const std::vector<int64_t> &ids, const std::vector<int64_t> &mask, size_t b, size_t L)
    {
        if (ids.size() != b * L || mask.size() != b * L) throw Error(SG_ERR_DIM, "ids/mask shape mismatch");
        std::vector<float> out(b * dimension());
        check(sg_encoder_encode(h_, ids.data(), mask.data(), (uint32_t)b, (uint32_t)L, out.data()));
        return out;
    }

private:
    sg_encoder *h_ = nullptr;
};

struct Surfaced { uint32_t candidate; float score; };

/// fuse_scores_full + recency/graph boosts + sort + `>= min_score` + truncate over SoA candidate features
inline std::vector<Surfaced> fuse_and_rank(const sg_candidate_feats &feats, uint32_t n, int64_t now_unix_s,
                                           const sg_fuse_weights *w = nullptr, const sg_fuse_config *cfg = nullptr, int device = 0)
{
    sg_fuse_weights dw; sg_fuse_default_weights(&dw);
    sg_fuse_config dc; sg_fuse_default_config(&dc);
    if (!w) w = &dw;
    if (!cfg) cfg = &dc;
    std::vector<uint32_t> order(cfg->max_results ? cfg->max_results : 1);
    std::vector<float> score(order.size());
    uint32_t m = 0;
    check(sg_fuse_topk(device, w, cfg, &feats, n, now_unix_s, order.data(), score.data(), &m, nullptr));
    std::vector<Surfaced> out;
    for (uint32_t i = 0; i < m; ++i) out.push_back({order[i], score[i]});
    return out;
}

struct MemoryId { uint64_t hi, lo; };   // Uuid as two big-endian halves

/// RRFusion (src/memory/hybrid_search.rs:529-594): weights are normalised inside the library like RRFusion::new
class RRFusion {
public:
    RRFusion(float k, std::vector<float> weights, int device = 0) : k_(k), weights_(std::move(weights)), device_(device) {}
    /// Each input list is in rank order; returns (MemoryId, rrf_score) sorted (score desc, MemoryId asc).
    std::vector<std::pair<MemoryId, float>> fuse(const std::vector<std::vector<MemoryId>> &ranked_lists) const
    {
        std::vector<uint32_t> offsets(ranked_lists.size() + 1, 0);
        std::vector<uint64_t> ids;
        for (size_t l = 0; l < ranked_lists.size(); ++l) {
            for (const MemoryId &m : ranked_lists[l]) { ids.push_back(m.hi); ids.push_back(m.lo); }
            offsets[l + 1] = (uint32_t)(ids.size() / 2);
        }
        const size_t n = ids.size() / 2;
        std::vector<uint64_t> out(2 * (n ? n : 1));
        std::vector<float> score(n ? n : 1);
        uint32_t m = 0;
        check(sg_rrf_fuse(device_, k_, weights_.data(), (uint32_t)weights_.size(), ids.data(), offsets.data(),
                          (uint32_t)ranked_lists.size(), out.data(), score.data(), &m));
        std::vector<std::pair<MemoryId, float>> r;
        for (uint32_t i = 0; i < m; ++i) r.push_back({MemoryId{out[2 * i], out[2 * i + 1]}, score[i]});
        return r;
    }
private:
    float k_;
    std::vector<float> weights_;
    int device_;
};

}  // namespace shodh
