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


def test_pair_kernel_large_with_prepass(sg, oracle, monkeypatch):
    """2M rows: the pre-pass / threshold seeding is active; compare with the exact path."""
    n, b = 2_000_000, 256
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(SEED, 0, n)
    queries, _ = oracle.synth_queries(SEED, n, b, 384)
    tc = st.search_batch(queries, 10)
    assert st.tc_stats() == (b, 0)
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
