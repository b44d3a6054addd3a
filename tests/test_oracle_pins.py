"""Pins the CPU oracle against every known-answer unit test the reference holds
for the hot-path functions, plus an independent
step-by-step numpy-f32 restatement of the two summation orders.

Each test names the reference test it ports (paths relative to /root/reference).
"""
import math

import numpy as np
import pytest

F = np.float32


# ---- independent restatements (pure python / numpy scalars), for cross-checking the C ----
def dot_scalar_py(a, b):
    """distance_inline.rs:157-173, one rounding per operation."""
    a = np.asarray(a, F); b = np.asarray(b, F)
    n = len(a); u = n & ~3
    s = F(0)
    for i in range(0, u, 4):
        p0, p1, p2, p3 = (F(a[i + j] * b[i + j]) for j in range(4))
        s = F(s + F(F(F(p0 + p1) + p2) + p3))
    for j in range(u, n):
        s = F(s + F(a[j] * b[j]))
    return float(s)


def fma32(x, y, z):
    # exact product and sum in float64 is not always a correct fma; use fractions for exactness
    from fractions import Fraction
    r = Fraction(float(x)) * Fraction(float(y)) + Fraction(float(z))
    return F(float(r)) if abs(r) < Fraction(2) ** 127 else F(float(r))


def dot_avx2_py(a, b):
    """distance_inline.rs:67-111 lane by lane."""
    a = np.asarray(a, F); b = np.asarray(b, F)
    n = len(a); simd = n & ~7
    lane = [F(0)] * 8
    i = 0
    while i + 16 <= simd:
        for j in range(8):
            lane[j] = fma32(a[i + j], b[i + j], lane[j])
        for j in range(8):
            lane[j] = fma32(a[i + 8 + j], b[i + 8 + j], lane[j])
        i += 16
    while i < simd:
        for j in range(8):
            lane[j] = fma32(a[i + j], b[i + j], lane[j])
        i += 8
    r = lane[0]
    for j in range(1, 8):
        r = F(r + lane[j])
    for j in range(simd, n):
        r = F(r + F(a[j] * b[j]))
    return float(r)


def test_fma32_helper_is_correctly_rounded():
    # float(Fraction) rounds to nearest double, then F() rounds again: guard against
    # double rounding by checking against math.fma when available (py3.13+) or known cases
    assert fma32(F(1.0), F(1.0), F(2.0 ** -30)) == F(1.0)
    assert fma32(F(3.0), F(5.0), F(-15.0)) == F(0.0)


@pytest.mark.parametrize("n", [0, 1, 3, 4, 7, 8, 9, 15, 16, 17, 31, 33, 384])
def test_c_dot_matches_python_restatement(oracle, n):
    rng = np.random.default_rng(100 + n)
    a = rng.standard_normal(n).astype(F)
    b = rng.standard_normal(n).astype(F)
    assert oracle.dot(a, b, oracle.VARIANT_SCALAR) == dot_scalar_py(a, b)
    assert oracle.dot(a, b, oracle.VARIANT_AVX2) == dot_avx2_py(a, b)


def euclid_py(a, b, avx2):
    """distance_inline.rs:284-305 (scalar) / :218-252 (avx2+fma), one rounding per operation."""
    a = np.asarray(a, F); b = np.asarray(b, F)
    n = len(a)
    if not avx2:
        s = F(0)
        for i in range(0, n & ~3, 4):
            d = [F(a[i + j] - b[i + j]) for j in range(4)]
            s = F(s + F(F(F(F(d[0] * d[0]) + F(d[1] * d[1])) + F(d[2] * d[2])) + F(d[3] * d[3])))
        for j in range(n & ~3, n):
            d = F(a[j] - b[j]); s = F(s + F(d * d))
        return float(s)
    lane = [F(0)] * 8
    for i in range(0, n & ~7, 8):
        for j in range(8):
            d = F(a[i + j] - b[i + j])
            lane[j] = fma32(d, d, lane[j])
    r = lane[0]
    for j in range(1, 8):
        r = F(r + lane[j])
    for j in range(n & ~7, n):
        d = F(a[j] - b[j]); r = F(r + F(d * d))
    return float(r)


@pytest.mark.parametrize("n", [4, 8, 13, 384])
def test_c_euclidean_norm_cosine_match_python_restatement(oracle, n):
    """both builds of euclidean_squared_inline / l2_norm_squared_inline / cosine_similarity_inline"""
    rng = np.random.default_rng(n)
    for _ in range(5):
        a = (rng.standard_normal(n) * 3).astype(F); b = (rng.standard_normal(n) * 0.7).astype(F)
        for variant in (0, 1):
            assert oracle.euclidean_squared(a, b, variant) == euclid_py(a, b, variant == 1)
            dot = dot_avx2_py if variant else dot_scalar_py
            assert oracle.l2_norm_squared(a, variant) == dot(a, a)       # same operation order as dot(a, a)
            den = F(np.sqrt(F(F(dot(a, a)) * F(dot(b, b)))))
            want = float(F(F(dot(a, b)) / den))
            assert oracle.cosine_similarity_inline(a, b, variant) == want
            assert oracle.distance(a, b, 2, variant) == float(F(F(1) - F(want)))
            assert oracle.distance(a, b, 1, variant) == oracle.euclidean_squared(a, b, variant)


def test_summation_orders_differ_somewhere(oracle):
    """The two reference builds are genuinely different functions; the oracle keeps both."""
    rng = np.random.default_rng(7)
    diff = 0
    for _ in range(50):
        a = rng.standard_normal(384).astype(F); b = rng.standard_normal(384).astype(F)
        diff += oracle.dot(a, b, 0) != oracle.dot(a, b, 1)
    assert diff > 0


# ---- distance_inline.rs:508-636 ----
def test_dot_product_correctness(oracle):
    """distance_inline.rs:515-529: [1..9].[9..1] == 165."""
    a = np.arange(1, 10, dtype=F); b = np.arange(9, 0, -1).astype(F)
    for v in (0, 1):
        assert abs(oracle.dot(a, b, v) - 165.0) < 1e-5


def test_euclidean_squared_correctness(oracle):
    """distance_inline.rs:531-545: ([1,2,3,4],[5,6,7,8]) == 64."""
    assert abs(oracle.euclidean_squared([1, 2, 3, 4], [5, 6, 7, 8]) - 64.0) < 1e-5


def test_l2_norm(oracle):
    """distance_inline.rs:547-556: |[3,4]| == 5."""
    assert abs(oracle.l2_norm([3, 4]) - 5.0) < 1e-5


def test_cosine_similarity_inline(oracle):
    """distance_inline.rs:558-588: same / opposite / orthogonal."""
    assert abs(oracle.cosine_similarity_inline([1, 0], [2, 0]) - 1.0) < 1e-5
    assert abs(oracle.cosine_similarity_inline([1, 0], [-1, 0]) + 1.0) < 1e-5
    assert abs(oracle.cosine_similarity_inline([1, 0], [0, 1])) < 1e-5


def test_is_normalized_after_normalize_inplace(oracle):
    """distance_inline.rs:590-597."""
    v = oracle.normalize([3, 4])
    assert abs(float(np.dot(v, v)) - 1.0) < 0.01


def test_normalized_distance(oracle):
    """distance_inline.rs:599-611: -dot ordering."""
    a = [1.0, 0.0]
    d = oracle.normalized_distance(a, [0.6, 0.8])
    assert d < 0
    assert oracle.normalized_distance(a, [0.8, 0.6]) < d


def test_large_vectors(oracle):
    """distance_inline.rs:613-635: 384-d ramps vs the naive iterator sum, tol 0.01."""
    a = (np.arange(384, dtype=F) * F(0.01)).astype(F)
    b = ((384 - np.arange(384)).astype(F) * F(0.01)).astype(F)
    naive = F(0)
    for x, y in zip(a, b):
        naive = F(naive + F(x * y))
    for v in (0, 1):
        assert abs(oracle.dot(a, b, v) - float(naive)) < 0.01
    naive_e = F(0)
    for x, y in zip(a, b):
        naive_e = F(naive_e + F(F(x - y) * F(x - y)))
    assert abs(oracle.euclidean_squared(a, b) - float(naive_e)) < 0.01


# ---- similarity.rs:50-124 ----
def test_cosine_similarity(oracle):
    """similarity.rs:55-67."""
    assert abs(oracle.cosine_similarity([1, 0, 0], [1, 0, 0]) - 1.0) < 1e-3
    assert abs(oracle.cosine_similarity([1, 0, 0], [0, 1, 0])) < 1e-3
    assert abs(oracle.cosine_similarity([1, 1], [1, 1]) - 1.0) < 1e-3


def test_cosine_similarity_edge_cases(oracle):
    """similarity.rs:69-81: length mismatch -> 0, zero vector -> 0."""
    assert oracle.cosine_similarity([1, 2], [1, 2, 3]) == 0.0
    assert oracle.cosine_similarity([0, 0, 0], [1, 2, 3]) == 0.0
    assert oracle.cosine_similarity([1, 2, 3], [0, 0, 0]) == 0.0


def test_cosine_similarity_negative_correlation(oracle):
    """similarity.rs:83-89."""
    assert abs(oracle.cosine_similarity([1, -1], [-1, 1]) + 1.0) < 1e-3


def test_top_k_similar_orders_and_limits(oracle):
    """similarity.rs:91-106."""
    cands = [[1, 0], [0.7, 0.7], [0, 1], [-1, 0]]
    pos, sc = oracle.top_k_similar([1, 0], cands, 2)
    assert list(pos) == [0, 1]
    assert sc[0] >= sc[1]


def test_top_k_similar_k_larger_than_candidates(oracle):
    """similarity.rs:108-115."""
    pos, _ = oracle.top_k_similar([1, 0], [[1, 0], [0, 1]], 10)
    assert len(pos) == 2


def test_top_k_similar_empty_candidates(oracle):
    """similarity.rs:117-123."""
    pos, _ = oracle.top_k_similar([1, 0], np.zeros((0, 2), F), 3)
    assert len(pos) == 0


def test_top_k_similar_is_stable_on_ties(oracle):
    """similarity.rs:40: stable sort on score only -> ties keep input order."""
    cands = [[0, 1], [2, 0], [1, 0], [3, 0]]
    pos, _ = oracle.top_k_similar([1, 0], cands, 4)
    assert list(pos) == [1, 2, 3, 0]


# ---- vamana.rs:1681-1711 (search) through the exact path (vamana.rs:775-777,1167-1188) ----
def test_vamana_build_and_search_exact(oracle):
    """vamana.rs:1697-1711: 5 vectors in R^4, query [0.9,0.1,0,0] -> 2 results, top-1 id 0."""
    vecs = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [0.5, 0.5, 0, 0]], F)
    ids, dist = oracle.brute_force_search(vecs, [0.9, 0.1, 0, 0], 2)
    assert len(ids) == 2 and ids[0] == 0
    assert dist[0] == F(-0.9)
    ids2, dist2 = oracle.brute_force_search(vecs, [0.9, 0.1, 0, 0], 2, as_written=True)
    assert list(ids2) == list(ids) and list(dist2) == list(dist)


def test_brute_force_order_ties_and_deleted(oracle):
    """vamana.rs:1175-1177 (deleted skipped), :1184 (dist total_cmp asc, id asc), truncate."""
    vecs = np.array([[1, 0], [0, 1], [1, 0], [-1, 0], [1, 0]], F)
    ids, dist = oracle.brute_force_search(vecs, [1, 0], 10)
    assert list(ids) == [0, 2, 4, 1, 3]
    assert list(dist) == [-1.0, -1.0, -1.0, -0.0, 1.0]
    ids, _ = oracle.brute_force_search(vecs, [1, 0], 10, deleted=[2])
    assert list(ids) == [0, 4, 1, 3]
    ids, _ = oracle.brute_force_search(vecs, [1, 0], 2, deleted=[0])
    assert list(ids) == [2, 4]


def test_total_order_key(oracle):
    """f32::total_cmp: -0.0 < +0.0; +NaN last; -NaN first."""
    k = oracle.total_order_key
    vals = [float("-inf"), -1.0, -0.0, 0.0, 1.0, float("inf")]
    keys = [k(v) for v in vals]
    assert keys == sorted(keys) and len(set(keys)) == len(keys)
    assert k(float("nan")) > k(float("inf"))


def test_neg_zero_sorts_before_pos_zero(oracle):
    # dot == +0 -> dist == -0 ; dot == -0 -> dist == +0
    vecs = np.array([[0, 1], [0, -1]], F)   # dots with [1,0]: +0*.. -> 0 ; second also 0 (1*0 + 0*-1 = -0+0)
    ids, dist = oracle.brute_force_search(vecs, [1, 0], 2)
    assert sorted(ids) == [0, 1]


# ---- retrieval.rs:2430-2438 test_vector family ----
from helpers import ref_test_vector  # retrieval.rs:2430-2438


@pytest.mark.parametrize("variant", [0, 1])
def test_test_vector_self_retrieval(oracle, variant):
    """retrieval.rs:2482-2529: every test_vector must retrieve its own id first."""
    dim, n = 384, 200
    corpus = np.stack([ref_test_vector(i, dim) for i in range(n)])
    for i in (0, 1, 57, 199):
        ids, dist = oracle.brute_force_search(corpus, corpus[i], 5, variant=variant)
        assert ids[0] == i
        assert abs(-dist[0] - 1.0) < 1e-5
        if i + 1 < n:
            assert abs(-dist[1] - 0.8) < 0.05  # "cosine ~0.8" per the reference doc comment


# ---- retrieval.rs:927-961 ----
def test_search_ids_post(oracle):
    vec_ids = [5, 3, 9, 1, 7]
    dists = [-0.9, -0.8, -0.8, -0.5, -0.95]
    vec2mem = np.full((10, 2), np.iinfo(np.uint64).max, np.uint64)
    vec2mem[5] = (0, 20); vec2mem[3] = (0, 10); vec2mem[9] = (0, 5); vec2mem[7] = (0, 20)  # 5,7 same memory
    mem, sim = oracle.search_ids_post(vec_ids, dists, vec2mem, 10)     # id 1 unmapped -> dropped
    assert [int(m[1]) for m in mem] == [20, 5, 10]                     # 0.95, then tie 0.8 by MemoryId asc
    assert list(sim) == [F(0.95), F(0.8), F(0.8)]
    mem, _ = oracle.search_ids_post(vec_ids, dists, vec2mem, 2)
    assert len(mem) == 2


# ---- relevance.rs:1790-2067 ----
def test_default_weights_sum_to_one(oracle):
    """relevance.rs:1930-1943."""
    assert abs(float(oracle.DEFAULT_WEIGHTS.sum()) - 1.0) < 1e-3


def test_calibrate_score(oracle):
    """relevance.rs:1945-1958."""
    assert oracle.calibrate_score(0.9) > 0.9
    assert oracle.calibrate_score(0.1) < 0.1
    assert abs(oracle.calibrate_score(0.5) - 0.5) < 1e-3


def test_score_fusion(oracle):
    """relevance.rs:1960-1979."""
    f = oracle.fuse_scores_full
    assert f(0.9, 0.9, 0.9, 0.9, 0.9, 16, 0.9) > 0.8
    assert f(0.1, 0.1, 0.1, 0.1, -0.9, 0, 0.1) < 0.3
    assert 0.2 < f(0.9, 0.1, 0.5, 0.7, 0.0, 2, 0.5) < 0.8
    assert f(0.9, 0.9, 0.9, 0.9, 0.0, 0, 0.5) > 0.5   # legacy fuse_scores = neutral momentum/access/graph


def test_fuse_scores_nan_inf_guard(oracle):
    """relevance.rs:2044-2066."""
    f = oracle.fuse_scores_full
    assert math.isfinite(f(float("nan"), 0.5, 0.5, 0.5, 0.0, 1, 0.5))
    assert math.isfinite(f(0.5, float("inf"), 0.5, 0.5, 0.0, 1, 0.5))
    assert math.isfinite(f(0.5, 0.5, float("-inf"), 0.5, 0.0, 1, 0.5))
    r = f(0.8, 0.5, 0.3, 0.6, 0.2, 3, 0.4)
    assert math.isfinite(r) and r > 0


def test_fuse_matches_formula(oracle):
    """Independent float64 evaluation of relevance.rs:529-594 within 1e-6."""
    def sig(x):
        return 0.0 if not math.isfinite(x) else 1 / (1 + math.exp(-10 * (x - 0.5)))
    rng = np.random.default_rng(3)
    w = oracle.DEFAULT_WEIGHTS.astype(np.float64)
    for _ in range(200):
        sem, ent, tag, imp, gr = rng.random(5)
        mom = rng.uniform(-1, 1); acc = int(rng.integers(0, 70))
        m = (mom + 1) / 2
        m = min(m * 1.5, 1) if m > 0.65 else (max(m * 0.3, 0) if m < 0.40 else m)
        a = 0.0 if acc == 0 else min(math.log2(acc + 1) / 4, 1)
        want = (w * np.array([sig(sem), sig(ent), sig(tag), sig(imp), sig(m), sig(a), sig(gr)])).sum()
        got = oracle.fuse_scores_full(F(sem), F(ent), F(tag), F(imp), F(mom), acc, F(gr))
        assert abs(got - want) < 2e-6


def test_recency_boost(oracle):
    """relevance.rs:1524-1547: linear decay within 24 h, cap 1.0, integer hours."""
    now = 1_000_000_000
    assert oracle.apply_recency_boost(0.5, now, now) == float(F(F(0.5) * F(1.2)))
    assert oracle.apply_recency_boost(0.5, now, now - 3599) == float(F(F(0.5) * F(1.2)))  # 0 whole hours
    assert oracle.apply_recency_boost(0.5, now, now - 25 * 3600) == 0.5
    assert oracle.apply_recency_boost(0.95, now, now) == 1.0
    assert oracle.apply_recency_boost(0.5, now, now + 7200) == 0.5     # future: negative -> huge u64
    h12 = oracle.apply_recency_boost(0.5, now, now - 12 * 3600)
    assert abs(h12 - 0.5 * 1.1) < 1e-6


def test_fuse_rank_sort_threshold_truncate(oracle):
    """relevance.rs:905-918: sort (score desc, created desc, id asc), >= 0.25, truncate."""
    now = 2_000_000_000
    n = 6
    feats = dict(semantic=[0.9, 0.9, 0.1, 0.9, 0.5, 0.9], entity=[0.9, 0.9, 0.0, 0.9, 0.0, 0.9],
                 tag=[0.5] * n, importance=[0.5] * n, momentum=[0.5, 0.5, -0.9, 0.5, 0.0, 0.5],
                 access_count=[4, 4, 0, 4, 1, 4], graph=[0.5, 0.5, 0.1, 0.5, 0.5, 0.5],
                 created_at=[now - 100 * 3600, now - 90 * 3600, now - 100 * 3600, now - 100 * 3600,
                             now - 100 * 3600, now - 90 * 3600])
    order, scores, allsc = oracle.fuse_rank(feats, now, max_results=10)
    # candidates 0,1,3,5 have identical features -> identical scores; tie-break created desc then id asc
    assert list(order[:4]) == [1, 5, 0, 3]
    assert all(s >= 0.25 for s in scores)
    assert 2 not in order          # all-low candidate falls under MIN_RELEVANCE_SCORE
    order2, _, _ = oracle.fuse_rank(feats, now, max_results=2)
    assert list(order2) == [1, 5]


# ---- minilm.rs:959-978 + :846-878 ----
def test_pool_finalize(oracle):
    rng = np.random.default_rng(5)
    h = rng.standard_normal((8, 384)).astype(F)
    att = np.array([1, 1, 1, 1, 1, 0, 0, 0])
    out = oracle.pool_finalize(h, att)
    want = h[:5].astype(np.float64).mean(0); want /= np.linalg.norm(want)
    assert np.allclose(out, want, atol=1e-6)
    assert abs(float(np.linalg.norm(out)) - 1.0) < 1e-5          # minilm.rs:1416-1417
    assert not oracle.pool_finalize(h, np.zeros(8, np.int64)).any()  # all-masked row -> zeros
    h2 = h.copy(); h2[0, 3] = np.nan; h2[1, 4] = np.inf
    out2 = oracle.pool_finalize(h2, att)
    assert np.isfinite(out2).all() and out2[3] == 0 and out2[4] == 0


# ---- synthetic generator sanity ----
def test_synth_is_deterministic_unit_norm_and_shardable(oracle):
    a = oracle.synth_corpus(23821, 0, 64)
    b = oracle.synth_corpus(23821, 32, 32)
    assert np.array_equal(a[32:], b)
    assert np.allclose(np.linalg.norm(a.astype(np.float64), axis=1), 1.0, atol=1e-6)
    h = oracle.synth_corpus(23821, 0, 64, f16_round=True)
    assert np.array_equal(h, a.astype(np.float16).astype(F))
    q, near = oracle.synth_queries(23821, 64, 8)
    assert np.allclose(np.linalg.norm(q.astype(np.float64), axis=1), 1.0, atol=1e-6)
    for i in range(1, 8, 2):
        assert float(q[i] @ a[int(near[i])]) > 0.99


# ---- hybrid_search.rs:969-1052, :1172-1240 (RRFusion unit tests) ----
def _mid(rng):
    return rng.integers(0, 2**63, 2, dtype=np.uint64)


def test_rrf_fusion_basic_disjoint_ties(oracle):
    rng = np.random.default_rng(11)
    id1, id2, id3 = _mid(rng), _mid(rng), _mid(rng)
    ids, sc = oracle.rrf_fuse(60.0, [0.5, 0.5], [np.stack([id1, id2, id3]), np.stack([id2, id1, id3])])
    assert len(ids) == 3 and np.array_equal(ids[2], id3)              # :1008
    by = {tuple(i): s for i, s in zip(ids, sc)}
    assert abs(by[tuple(id1)] - by[tuple(id2)]) < 1e-4 and by[tuple(id3)] < by[tuple(id1)]
    ids, sc = oracle.rrf_fuse(60.0, [0.5, 0.5], [id1[None], id2[None]])
    assert len(ids) == 2 and abs(sc[0] - sc[1]) < 1e-3                # :1026-1028
    lo, hi = np.zeros(2, np.uint64), np.full(2, 2**64 - 1, np.uint64)  # :1038-1051 tie -> lower MemoryId first
    ids, sc = oracle.rrf_fuse(60.0, [0.5, 0.5], [hi[None], lo[None]])
    assert sc[0] == sc[1] and np.array_equal(ids[0], lo) and np.array_equal(ids[1], hi)


def test_rrf_weights_and_k(oracle):
    rng = np.random.default_rng(12)
    id1, id2, id3 = _mid(rng), _mid(rng), _mid(rng)
    bm25, vec = np.stack([id1, id2]), np.stack([id2, id1])
    assert np.array_equal(oracle.rrf_fuse(60.0, [0.8, 0.2], [bm25, vec])[0][0], id1)   # :1187-1188
    assert np.array_equal(oracle.rrf_fuse(60.0, [0.2, 0.8], [bm25, vec])[0][0], id2)   # :1191-1195
    l1, l2 = np.stack([id1, id2, id3]), np.stack([id3, id2, id1])
    rel = []
    for k in (1.0, 100.0):                                                              # :1198-1238
        ids, sc = oracle.rrf_fuse(k, [0.5, 0.5], [l1, l2])
        s2 = sc[[np.array_equal(i, id2) for i in ids]][0]
        rel.append(s2 / sc[0])
    assert rel[1] >= rel[0] - 0.01
    # the formula itself: w/(k + rank + 1), weights normalised, f32
    ids, sc = oracle.rrf_fuse(45.0, [0.35, 0.40], [id1[None], id2[None]])
    w = np.float32(0.35) + np.float32(0.40)
    want = sorted([np.float32(0.35) / w / np.float32(46.0), np.float32(0.40) / w / np.float32(46.0)], reverse=True)
    assert list(sc) == want
