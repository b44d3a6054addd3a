// Exercises the C++ host mirror (shodh_memory_b200/host/shodh_gpu.hpp) like the reference's own unit
// tests exercise VamanaIndex: vamana.rs:1697-1711 (5 vectors, top-1), retrieval.rs:2430-2438 (self
// retrieval), similarity.rs:108-115 (k > candidates), vamana.rs:766-768 (empty -> Ok([])).
// Built and run by tests/test_host_cpp.py (GPU) / compile-checked on CPU.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../shodh_memory_b200/host/shodh_gpu.hpp"

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static std::vector<float> test_vector(size_t i, size_t dim)   // retrieval.rs:2430-2438
{
    std::vector<float> v(dim, 1.0f);
    v[i % dim] += 0.5f * std::sqrt((float)dim);
    float s = 0.0f;
    for (float x : v) s += x * x;
    float norm = std::sqrt(s);
    for (float &x : v) x /= norm;
    return v;
}

int main()
{
    try {
        shodh::GpuVectorStore st(384, SG_DTYPE_F32, 0);
        std::vector<float> q0 = test_vector(0, 384);
        REQUIRE(st.search(q0, 5).empty());                        // empty index -> Ok(vec![])
        const size_t n = 300;
        for (size_t i = 0; i < n; ++i) REQUIRE(st.add_vector(test_vector(i, 384)) == i);   // dense sequential ids
        REQUIRE(st.len() == n);
        for (size_t i : {0u, 7u, 299u}) {
            auto r = st.search(test_vector(i, 384), 5);
            REQUIRE(r.size() == 5 && r[0].first == i);
            REQUIRE(std::fabs(-r[0].second - 1.0f) < 1e-5f);
            for (size_t j = 1; j < r.size(); ++j)                  // ascending (dist, id)
                REQUIRE(r[j - 1].second < r[j].second || (r[j - 1].second == r[j].second && r[j - 1].first < r[j].first));
        }
        REQUIRE(st.mark_deleted(7) && !st.mark_deleted(100000) && st.is_deleted(7) && st.deleted_count() == 1);
        REQUIRE(st.search(test_vector(7, 384), 3)[0].first != 7);
        REQUIRE(st.search(q0, 1000).size() == n - 1);              // k > live rows
        std::vector<float> batch;
        for (size_t i = 0; i < 9; ++i) { auto v = test_vector(10 + i, 384); batch.insert(batch.end(), v.begin(), v.end()); }
        auto rb = st.search_batch(batch, 9, 3);
        for (size_t i = 0; i < 9; ++i) REQUIRE(rb[i][0].first == 10 + i);
        bool threw = false;
        try { st.add_vector(std::vector<float>(10, 0.0f)); } catch (const shodh::Error &e) { threw = e.code == SG_ERR_DIM; }
        REQUIRE(threw);
        // relevance ranking through the mirror (relevance.rs:1960-1979 pins)
        float sem[2] = {0.9f, 0.1f}, ent[2] = {0.9f, 0.1f}, tag[2] = {0.9f, 0.1f}, imp[2] = {0.9f, 0.1f}, mom[2] = {0.9f, -0.9f}, gr[2] = {0.9f, 0.1f};
        uint32_t acc[2] = {16, 0};
        int64_t created[2] = {1000, 1000};
        sg_candidate_feats f{sem, ent, tag, imp, mom, acc, gr, created};
        auto ranked = shodh::fuse_and_rank(f, 2, 1000 + 100 * 3600);
        REQUIRE(ranked.size() == 1 && ranked[0].candidate == 0 && ranked[0].score > 0.8f);
        // similarity.rs:89-105: top_k_similar orders by cosine and truncates
        {
            std::vector<float> cands(3 * 32, 0.0f), q(32, 0.0f);
            cands[0 * 32 + 1] = 1.0f;                              // orthogonal
            cands[1 * 32 + 0] = 1.0f;                              // identical
            cands[2 * 32 + 0] = -1.0f;                             // opposite
            q[0] = 1.0f;
            auto top = shodh::top_k_similar(q, cands, 32, 2);
            REQUIRE(top.size() == 2 && top[0].second == 1 && top[0].first == 1.0f && top[1].second == 0 && top[1].first == 0.0f);
        }
        // the graph arm: "not built" is an error, a built graph answers like vamana.rs:1697-1711
        {
            bool nb = false;
            try { st.search_ann(q0, 3); } catch (const shodh::Error &e) { nb = e.code == SG_ERR_NOT_BUILT; }
            REQUIRE(nb);
            st.build_graph(16, 40, 1.2f, 1);
            auto r = st.search_ann(test_vector(11, 384), 5);
            REQUIRE(r.size() == 5 && r[0].first == 11);
            REQUIRE(st.add_vector(test_vector(300, 384)) == 300);  // linked into the graph like add_vector
            REQUIRE(st.search_ann(test_vector(300, 384), 3).size() == 3);
        }
        // storage: the rows that were appended one by one fit, and the matrix is reported consistently
        {
            auto mi = st.memory_info();
            REQUIRE(mi.capacity_rows >= st.len() && mi.bytes >= st.len() * 384 * 4 && mi.chunks >= 1);
        }
        // pq.rs:622-641 (test_encode_decode shape): a quantizer with one centroid per axis point encodes to the nearest
        // centroid, the compressed store returns (id, distance) ascending
        {
            const uint32_t dim = 32, ds = 8, kc = 4, m = dim / ds;
            std::vector<float> cent((size_t)m * kc * ds);
            for (uint32_t s2 = 0; s2 < m; ++s2)
                for (uint32_t c = 0; c < kc; ++c)
                    for (uint32_t d = 0; d < ds; ++d) cent[((size_t)s2 * kc + c) * ds + d] = (float)c;   // centroid c = (c, c, ..., c)
            shodh::GpuProductQuantizer pq(cent, dim, ds, kc, 0);
            std::vector<float> v(dim);
            for (uint32_t d = 0; d < dim; ++d) v[d] = (float)((d / ds) % kc) + 0.1f;
            auto codes = pq.encode(v);
            REQUIRE(codes.size() == m);
            for (uint32_t s2 = 0; s2 < m; ++s2) REQUIRE(codes[s2] == s2 % kc);
            pq.add(42, v);
            std::vector<float> w(dim, 3.0f);
            pq.add(7, w);
            REQUIRE(pq.len() == 2);
            auto r = pq.search(v, 2);
            REQUIRE(r.size() == 2 && r[0].first == 42 && r[1].first == 7 && r[0].second <= r[1].second);
        }
        std::printf("HOST_MIRROR_OK\n");
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
}
