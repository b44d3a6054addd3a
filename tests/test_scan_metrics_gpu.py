"""GPU parity for the metrics beside NormalizedDotProduct.

VamanaIndex::distance (vamana.rs:755-761) dispatches to euclidean_squared_inline and
1 - cosine_similarity_inline; similarity.rs::top_k_similar ranks by cosine_similarity.  The exact
sweep scores them in the reference's summation order, so ids and f32 bit patterns must equal the
CPU oracle's.  Everything goes through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F = np.float32
SEED = 23821
METRICS = {"euclidean": 1, "cosine": 2}


def bits(a):
    return np.ascontiguousarray(a, F).view(np.uint32)


def unnormalised(rng, n, dim):
    """rows with norms spread over [0.2, 5] -- Euclidean / Cosine accept any vectors."""
    x = rng.standard_normal((n, dim)).astype(F)
    return (x * rng.uniform(0.2, 5.0, (n, 1)).astype(F)).astype(F)


@pytest.mark.parametrize("order,variant", [("scalar", 0), ("avx2", 1)])
@pytest.mark.parametrize("dtype", ["f16", "f32"])
@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
def test_metric_bit_identical(sg, oracle, metric, dtype, order, variant):
    rng = np.random.default_rng(7)
    n, dim = 5000, 384
    corpus = unnormalised(rng, n, dim)
    seen = oracle.round_f16(corpus) if dtype == "f16" else corpus
    st = sg.GpuVectorStore(dim, dtype, metric=metric)
    st.set_order(order)
    st.add_vectors(corpus)
    st.mark_deleted(3)
    for b in (1, 2, 3, 4, 7):
        queries = unnormalised(rng, b, dim)
        queries[0] = corpus[17] * F(1.5)            # a planted neighbour (cosine 1, Euclidean far from 0)
        for k in (1, 10, 100):
            ids, dist, counts = st.search_batch(queries, k)
            for q in range(b):
                rid, rdist = oracle.brute_force_search(seen, queries[q], k, deleted=[3], metric=METRICS[metric],
                                                       variant=variant)
                c = int(counts[q])
                assert c == len(rid)
                assert np.array_equal(ids[q, :c], rid), (metric, dtype, order, b, k, q, ids[q, :c], rid)
                assert np.array_equal(bits(dist[q, :c]), bits(rdist)), (metric, dtype, order, b, k, q)
    st.close()


def test_reference_unit_pins(sg):
    """distance_inline.rs:515-635 known answers through the device path (dim padded to 32 with zeros)."""
    def pad(v):
        out = np.zeros(32, F)
        out[:len(v)] = v
        return out
    st = sg.GpuVectorStore(32, "f32", metric="euclidean")
    st.add_vectors(np.stack([pad([5, 6, 7, 8]), pad([1, 2, 3, 4])]))
    r = st.search(pad([1, 2, 3, 4]), 2)
    assert r == [(1, 0.0), (0, 64.0)]               # test_euclidean_squared: |[1,2,3,4]-[5,6,7,8]|^2 = 64
    st.close()
    st = sg.GpuVectorStore(32, "f32", metric="cosine")
    st.add_vectors(np.stack([pad([1, 2, 3]), pad([-1, -2, -3]), pad([0, 1, 0]), pad([0, 0, 0])]))
    r = dict(st.search(pad([1, 2, 3]), 4))
    assert abs(r[0]) < 1e-6 and abs(r[1] - 2.0) < 1e-6          # same / opposite direction
    assert r[3] == 1.0                                          # zero vector: cos 0 -> distance 1
    r = dict(st.search(pad([1, 0, 0]), 4))
    assert abs(r[2] - 1.0) < 1e-6                               # orthogonal
    st.close()


@pytest.mark.parametrize("order,variant", [("scalar", 0), ("avx2", 1)])
def test_top_k_similar_matches_oracle(sg, oracle, order, variant):
    rng = np.random.default_rng(11)
    n, dim = 3000, 384
    cands = unnormalised(rng, n, dim)
    cands[100] = 0.0                                # zero norm -> similarity 0.0
    cands[200] = cands[50]                          # exact tie: the stable sort keeps position order
    cands[201] = cands[50]
    query = (cands[50] + F(0.05) * rng.standard_normal(dim)).astype(F)
    for k in (1, 5, 64, 1024):
        pos, sc = sg.top_k_similar(query, cands, k, order=order)
        rpos, rsc = oracle.top_k_similar(query, cands, k, variant=variant)
        assert np.array_equal(pos, rpos), (k, pos[:8], rpos[:8])
        assert np.array_equal(sc, rsc) and np.array_equal(bits(sc) & 0x7FFFFFFF, bits(rsc) & 0x7FFFFFFF)
    # the planted duplicates tie and come out in input order
    pos, sc = sg.top_k_similar(cands[50], cands, 3, order=order)
    assert list(pos) == [50, 200, 201] and sc[0] == sc[1] == sc[2]


def test_top_k_similar_edge_cases(sg, oracle):
    """similarity.rs:55-123: identical -> 1, orthogonal -> 0, opposite -> -1, length mismatch -> 0, empty."""
    e = np.eye(32, dtype=F)
    pos, sc = sg.top_k_similar(e[0], np.stack([e[1], e[0], -e[0]]), 3)
    assert list(pos) == [1, 0, 2] and list(sc) == [1.0, 0.0, -1.0]
    assert abs(sg.cosine_similarity(e[0] + e[1], e[0]) - 1 / np.sqrt(2)) < 1e-6
    pos, sc = sg.top_k_similar(np.ones(64, F), np.ones((4, 32), F), 2)          # len(a) != len(b) -> 0.0
    assert list(pos) == [0, 1] and list(sc) == [0.0, 0.0]
    pos, sc = sg.top_k_similar(e[0], np.zeros((0, 32), F), 5)
    assert len(pos) == 0
    # all-zero query: every score 0 (norm_a == 0), input order kept
    pos, sc = sg.top_k_similar(np.zeros(32, F), e[:5], 3)
    assert list(pos) == [0, 1, 2] and not sc.any()
    # NaN query: every score NaN; OrderedFloat treats NaNs as equal -> input order
    q = e[0].copy()
    q[3] = np.nan
    pos, sc = sg.top_k_similar(q, e[:5], 3)
    rpos, rsc = oracle.top_k_similar(q, e[:5], 3)
    assert np.array_equal(pos, rpos) and np.isnan(sc).all() and np.isnan(rsc).all()


def test_append_from_device_memory(sg, oracle):
    """sg_store_append_device: rows that are already in HBM (encoder output) go in without a host round trip"""
    import torch
    rows = oracle.synth_corpus(SEED, 0, 3000, 384)
    d = torch.from_numpy(rows).cuda()
    for dtype in ("f16", "f32"):
        st = sg.GpuVectorStore(384, dtype)
        assert st.add_vectors(rows[:1000]) == 0
        assert st.add_vectors_device(d[1000:].data_ptr(), 2000, torch.cuda.current_stream().cuda_stream) == 1000
        want = oracle.round_f16(rows) if dtype == "f16" else rows
        assert len(st) == 3000 and np.array_equal(bits(st.extract_all_vectors()), bits(want))
        q, _ = oracle.synth_queries(SEED, 3000, 6, 384)
        ids, dist, _ = st.search_batch(q, 10)
        rid, rdist, _, _ = oracle.brute_force_search_batch(want, q, 10)
        assert np.array_equal(ids, rid) and np.array_equal(bits(dist), bits(rdist))
        st.close()
