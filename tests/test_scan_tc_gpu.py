"""GPU parity for the tensor-core sweep (path B): batched queries must return EXACTLY what the
exact sweep / the CPU oracle return (ids and f32 distance bit patterns), because every result is
re-scored in the reference's summation order and certified -- or re-run on the exact path."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F = np.float32
SEED = 23821


def bits(a):
    return np.ascontiguousarray(a, F).view(np.uint32)


def check_batch(oracle, store, seen, queries, k, deleted=None, variant=0):
    ids, dist, counts = store.search_batch(queries, k)
    rid, rdist, rcounts, _ = oracle.brute_force_search_batch(seen, queries, k, deleted=deleted, variant=variant)
    assert np.array_equal(counts, rcounts)
    for q in range(queries.shape[0]):
        c = int(counts[q])
        assert np.array_equal(ids[q, :c], rid[q, :c]), (q, ids[q, :c], rid[q, :c])
        assert np.array_equal(bits(dist[q, :c]), bits(rdist[q, :c])), (q, dist[q, :c], rdist[q, :c])
    return ids, dist


@pytest.mark.parametrize("n", [4096, 5000, 100_000])
@pytest.mark.parametrize("b", [5, 64, 128, 129, 300])
def test_batched_matches_oracle(sg, oracle, n, b):
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(SEED, 0, n)
    seen = oracle.synth_corpus(SEED, 0, n, 384, f16_round=True)
    queries, near = oracle.synth_queries(SEED, n, b, 384)
    before = st.tc_stats()
    ids, _ = check_batch(oracle, st, seen, queries, 10)
    after = st.tc_stats()
    assert after[0] - before[0] == b                      # the tensor path really served this batch
    for q in range(1, b, 2):
        assert ids[q, 0] == near[q]
    st.close()


def test_batched_equals_exact_path_and_both_orders(sg, oracle):
    n, b = 200_000, 256
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(SEED, 0, n)
    queries, _ = oracle.synth_queries(SEED, n, b, 384)
    for order in ("scalar", "avx2"):
        st.set_order(order)
        tc = st.search_batch(queries, 10)
        os.environ["SG_TC_DISABLE"] = "1"
        try:
            ex = st.search_batch(queries, 10)
        finally:
            del os.environ["SG_TC_DISABLE"]
        for x, y in zip(tc, ex):
            assert x.tobytes() == y.tobytes()
    q_tot, flagged = st.tc_stats()
    assert q_tot == 2 * b and flagged <= b // 8            # certification almost always succeeds
    st.close()


def test_batched_with_deleted_rows_and_k_variants(sg, oracle):
    n, b = 50_000, 40
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(SEED, 0, n)
    seen = oracle.synth_corpus(SEED, 0, n, 384, f16_round=True)
    queries, near = oracle.synth_queries(SEED, n, b, 384)
    deleted = sorted({int(x) for x in near[1::2][:10]} | {0, 127, 128, 49_999})
    for d in deleted:
        st.mark_deleted(d)
    for k in (1, 10, 24):
        check_batch(oracle, st, seen, queries, k, deleted=deleted)
    st.close()


def test_duplicates_force_the_exact_rerun(sg, oracle):
    """60 identical rows tie exactly: more ties than the candidate slack, so certification must
    fail for queries aimed at them and the exact re-run must still give the id-ordered answer."""
    n = 20_000
    base = oracle.synth_corpus(SEED, 0, n, 384)
    base[1000:1060] = base[7]
    st = sg.GpuVectorStore(384, "f16")
    st.add_vectors(base)
    seen = oracle.round_f16(base)
    queries, _ = oracle.synth_queries(SEED, n, 16, 384)
    queries[3] = seen[7]
    queries[9] = seen[7]
    before = st.tc_stats()
    ids, _ = check_batch(oracle, st, seen, queries, 10)
    after = st.tc_stats()
    assert list(ids[3]) == [7] + list(range(1000, 1009))
    assert after[1] - before[1] >= 2                      # those two could not be certified
    st.close()


def test_batched_other_dims(sg, oracle):
    for dim in (64, 128, 256):
        n = 30_000
        corpus = oracle.synth_corpus(SEED, 0, n, dim)
        st = sg.GpuVectorStore(dim, "f16")
        st.add_vectors(corpus)
        queries, _ = oracle.synth_queries(SEED, n, 33, dim)
        before = st.tc_stats()[0]
        check_batch(oracle, st, oracle.round_f16(corpus), queries, 10)
        assert st.tc_stats()[0] - before == 33
        st.close()


def test_unnormalised_rows_still_exact(sg, oracle):
    """The certification bound scales with the largest row norm in the store."""
    rng = np.random.default_rng(4)
    n = 20_000
    corpus = (rng.standard_normal((n, 384)) * rng.uniform(0.1, 3.0, (n, 1))).astype(F)
    st = sg.GpuVectorStore(384, "f16")
    st.add_vectors(corpus)
    queries = rng.standard_normal((20, 384)).astype(F)
    check_batch(oracle, st, oracle.round_f16(corpus), queries, 10)
    st.close()


@pytest.mark.parametrize("n", [4096, 70_001])
@pytest.mark.parametrize("b", [5, 129, 256, 300])
def test_pair_kernel_matches_oracle(sg, oracle, n, b, monkeypatch):
    """CTA-pair form (tcgen05.mma.cta_group::2, 256 queries per sweep), forced for every batch size."""
    monkeypatch.setenv("SG_TC_PAIR", "2")
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(SEED, 0, n)
    seen = oracle.synth_corpus(SEED, 0, n, 384, f16_round=True)
    queries, near = oracle.synth_queries(SEED, n, b, 384)
    for d in (3, 128, 255, 256, n - 1):
        st.mark_deleted(d)
    before = st.tc_stats()
    check_batch(oracle, st, seen, queries, 10, deleted=[3, 128, 255, 256, n - 1])
    after = st.tc_stats()
    assert after[0] - before[0] == b
    assert after[1] - before[1] <= max(2, b // 8)          # a wrong row mapping would flag everything
    st.close()


def test_pair_kernel_large(sg, oracle, monkeypatch):
    """2M rows (528 tiles per CTA pair would be 10M; here 106): sample seeding + histogram tightening; compare with the exact path."""
    n, b = 2_000_000, 256
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(SEED, 0, n)
    queries, _ = oracle.synth_queries(SEED, n, b, 384)
    tc = st.search_batch(queries, 10)
    served, flagged = st.tc_stats()
    assert served == b and flagged <= 2
    monkeypatch.setenv("SG_TC_DISABLE", "1")
    ex = st.search_batch(queries[:16], 10)
    for x, y in zip(tc, ex):
        assert x[:16].tobytes() == y.tobytes()
    st.close()


def test_many_query_tiles(sg, oracle):
    """b = 600: three 256-query sweeps (pair kernel), the last one partially filled."""
    n, b = 30_000, 600
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(SEED, 0, n)
    seen = oracle.synth_corpus(SEED, 0, n, 384, f16_round=True)
    queries, _ = oracle.synth_queries(SEED, n, b, 384)
    check_batch(oracle, st, seen, queries, 10)
    assert st.tc_stats()[0] == b
    st.close()


@pytest.mark.parametrize("b,k", [(64, 25), (64, 120), (256, 64), (256, 120), (300, 128)])
def test_large_k_takes_the_tensor_path(sg, oracle, b, k):
    """The reference's own call shape is k = 4 * limit, limit = 3 * max_results: k = 120 for a top-10 recall
    (retrieval.rs:912-918, memory/mod.rs:3585).  Up to k = 128 the batch runs on the tensor-core sweep."""
    n = 100_000
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(SEED, 0, n)
    seen = oracle.synth_corpus(SEED, 0, n, 384, f16_round=True)
    queries, _ = oracle.synth_queries(SEED, n, b, 384)
    for d in (0, 31, 4095, n - 1):
        st.mark_deleted(d)
    before = st.tc_stats()
    check_batch(oracle, st, seen, queries, k, deleted=[0, 31, 4095, n - 1])
    after = st.tc_stats()
    assert after[0] - before[0] == b
    assert after[1] - before[1] <= max(2, b // 8)
    st.close()


def test_k_beyond_128_falls_back_to_the_exact_sweep(sg, oracle):
    n, b, k = 20_000, 8, 200
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(SEED, 0, n)
    seen = oracle.synth_corpus(SEED, 0, n, 384, f16_round=True)
    queries, _ = oracle.synth_queries(SEED, n, b, 384)
    check_batch(oracle, st, seen, queries, k)
    assert st.tc_stats()[0] == 0
    st.close()


def test_tc_mode_switch(sg, oracle):
    n, b = 50_000, 32
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(SEED, 0, n)
    queries, _ = oracle.synth_queries(SEED, n, b, 384)
    tc = st.search_batch(queries, 10)
    assert st.tc_stats()[0] == b
    st.set_tc_mode("exact")
    ex = st.search_batch(queries, 10)
    assert st.tc_stats()[0] == b                          # unchanged: path A served it
    st.set_tc_mode("auto")
    for x, y in zip(tc, ex):
        assert x.tobytes() == y.tobytes()
    st.close()


def test_rows_sorted_by_score_stay_exact(sg, oracle):
    """Adversarial row order: similarity to the queries' direction INCREASES with the row number, so every
    tile beats the running threshold and candidate regions overflow.  Whatever the sweep does (tighten through
    the shared histogram, or flag the queries for the exact re-run) the answers must be the oracle's."""
    rng = np.random.default_rng(11)
    n, dim = 60_000, 384
    d = rng.standard_normal(dim).astype(F)
    d /= np.linalg.norm(d)
    base = rng.standard_normal((n, dim)).astype(F)
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    t = np.linspace(0.0, 0.9, n, dtype=F)[:, None]
    rows = np.ascontiguousarray((1 - t) * base + t * d, F)
    st = sg.GpuVectorStore(dim, "f16")
    st.add_vectors(rows)
    seen = oracle.round_f16(rows)
    queries = np.ascontiguousarray(d[None, :] + 0.05 * rng.standard_normal((40, dim)).astype(F), F)
    check_batch(oracle, st, seen, queries, 10)
    # and the reverse order (best rows first: the sample already sees them)
    st2 = sg.GpuVectorStore(dim, "f16")
    st2.add_vectors(rows[::-1].copy())
    check_batch(oracle, st2, seen[::-1].copy(), queries, 10)
    st.close()
    st2.close()


def c_acc(dim):
    """the certificate's accumulation-error coefficient (csrc/scan_tc.cu::tc_c_acc)"""
    return 1.25 * (dim * 2.0 ** -23 + (dim // 4 + 4) * 2.0 ** -24)


@pytest.mark.parametrize("dim,n,b", [(384, 400_000, 256), (640, 24_000, 64)])
def test_accumulation_error_bound(sg, oracle, dim, n, b):
    """MEASURE max |tensor-core score - reference-order dot| over >= 1e8 (query, row) pairs (dim 384) and
    compare it with the bound the certificate assumes, c_acc(dim) * |q| * |row|.  Queries are fp16-representable,
    so the whole difference is accumulation order / rounding (the fp16 rounding of a real query is bounded
    separately and exactly by err_q * R)."""
    rng = np.random.default_rng(1)
    st = sg.GpuVectorStore(dim, "f16")
    if dim == 384:
        st.append_synthetic(SEED, 0, n)
        rnorm = 1.001
    else:
        rows = rng.standard_normal((n, dim)).astype(F) * rng.uniform(0.2, 2.0, (n, 1)).astype(F)
        st.add_vectors(rows)
        rnorm = float(np.linalg.norm(oracle.round_f16(rows).astype(np.float64), axis=1).max())
    q = rng.standard_normal((b, dim)).astype(F)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q = oracle.round_f16(q)
    worst = 0.0
    pairs = 0
    step = 100_000
    for lo in range(0, n, step):
        m = min(step, n - lo)
        tc, ex = st.debug_scores(q, lo, m)
        assert np.isfinite(tc).all() and np.isfinite(ex).all()
        worst = max(worst, float(np.abs(tc.astype(np.float64) - ex.astype(np.float64)).max()))
        pairs += tc.size
    qn = float(np.linalg.norm(q.astype(np.float64), axis=1).max())
    bound = c_acc(dim) * qn * rnorm
    print(f"dim {dim}: {pairs:.3g} pairs, max |tc - exact| = {worst:.3e}, bound c_acc*|q|*R = {bound:.3e}, "
          f"margin x{bound / max(worst, 1e-30):.1f}")
    assert pairs >= (1e8 if dim == 384 else 1e6)
    assert worst <= bound
    # spot-check the diagnostics themselves against the oracle
    corpus = oracle.synth_corpus(SEED, 0, 64, dim, f16_round=True) if dim == 384 else oracle.round_f16(rows[:64])
    tc, ex = st.debug_scores(q[:4], 0, 64)
    for i in range(4):
        ref = -oracle.score_rows(corpus, q[i], np.arange(64, dtype=np.uint32))
        assert np.array_equal(bits(ex[i]), bits(ref))
    st.close()


def test_large_dim_tensor_path(sg, oracle):
    """dim = 640 (the largest query block that leaves room for a 3-stage row ring; wider stores take the exact
    sweep): the certificate coefficient is derived from dim; results stay exact."""
    rng = np.random.default_rng(2)
    n, dim, b = 8192, 640, 24
    corpus = rng.standard_normal((n, dim)).astype(F)
    corpus /= np.linalg.norm(corpus, axis=1, keepdims=True)
    st = sg.GpuVectorStore(dim, "f16")
    st.add_vectors(corpus)
    queries = rng.standard_normal((b, dim)).astype(F)
    before = st.tc_stats()[0]
    check_batch(oracle, st, oracle.round_f16(corpus), queries, 10)
    assert st.tc_stats()[0] - before == b
    st.close()


def test_wide_rows_fall_back_to_the_exact_sweep(sg, oracle):
    """dim = 2048: the resident query block does not fit shared memory -> path A, still exact."""
    rng = np.random.default_rng(3)
    n, dim, b = 4096, 2048, 8
    corpus = rng.standard_normal((n, dim)).astype(F)
    st = sg.GpuVectorStore(dim, "f16")
    st.add_vectors(corpus)
    queries = rng.standard_normal((b, dim)).astype(F)
    check_batch(oracle, st, oracle.round_f16(corpus), queries, 10)
    assert st.tc_stats()[0] == 0
    st.close()
