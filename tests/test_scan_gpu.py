"""GPU parity tests for the exact scan + top-k (SURVEY.md 8a rows a5,a6,a9; Appendix C).

Everything goes through the C ABI (shodh_memory_b200.GpuVectorStore -> libshodh_gpu.so) and is
compared BIT-FOR-BIT (ids and f32 distance bit patterns) with the CPU oracle on the same inputs.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F = np.float32
SEED = 23821


def bits(a):
    return np.ascontiguousarray(a, F).view(np.uint32)


def assert_same(oracle, store, corpus_seen, queries, k, deleted=None, variant=0):
    ids, dist, counts = store.search_batch(queries, k)
    for q in range(queries.shape[0]):
        rid, rdist = oracle.brute_force_search(corpus_seen, queries[q], k, deleted=deleted, variant=variant)
        c = int(counts[q])
        assert c == len(rid), (q, c, len(rid))
        assert np.array_equal(ids[q, :c], rid), (q, ids[q, :c], rid)
        assert np.array_equal(bits(dist[q, :c]), bits(rdist)), (q, dist[q, :c], rdist)
        assert (ids[q, c:] == 0xFFFFFFFF).all() and np.isinf(dist[q, c:]).all()


def test_synthetic_generator_bit_identical(sg, oracle):
    for dtype in ("f16", "f32"):
        st = sg.GpuVectorStore(384, dtype)
        st.append_synthetic(SEED, 100, 777)
        want = oracle.synth_corpus(SEED, 100, 777, 384, f16_round=(dtype == "f16"))
        assert np.array_equal(bits(st.get_vectors(0, 777)), bits(want))
        st.close()


def test_append_rounds_like_numpy_f16(sg, oracle):
    rng = np.random.default_rng(1)
    rows = rng.standard_normal((1000, 128)).astype(F)
    st = sg.GpuVectorStore(128, "f16")
    assert st.add_vectors(rows[:600]) == 0
    assert st.add_vectors(rows[600:]) == 600          # ids dense, sequential (vamana.rs:854-855)
    assert len(st) == 1000
    got = st.extract_all_vectors()
    assert np.array_equal(bits(got), bits(rows.astype(np.float16).astype(F)))
    assert np.array_equal(bits(got), bits(oracle.round_f16(rows)))
    st.close()


@pytest.mark.parametrize("dtype", ["f16", "f32"])
@pytest.mark.parametrize("n", [1, 5, 31, 32, 33, 1000, 4097])
def test_small_corpora_all_batch_shapes(sg, oracle, dtype, n):
    corpus = oracle.synth_corpus(SEED, 0, n, 384)
    seen = oracle.round_f16(corpus) if dtype == "f16" else corpus
    st = sg.GpuVectorStore(384, dtype)
    st.add_vectors(corpus)
    for b in (1, 2, 3, 4, 5, 9):
        queries, _ = oracle.synth_queries(SEED + b, n, b, 384)
        for k in (1, 10, 100):
            assert_same(oracle, st, seen, queries, k)
    st.close()


@pytest.mark.parametrize("order,variant", [("scalar", 0), ("avx2", 1)])
@pytest.mark.parametrize("dtype", ["f16", "f32"])
def test_100k_both_summation_orders(sg, oracle, dtype, order, variant):
    """config[0]: 100k x 384, brute-force top-10 (and the reference's k=120)."""
    n = 100_000
    st = sg.GpuVectorStore(384, dtype)
    st.append_synthetic(SEED, 0, n)
    st.set_order(order)
    seen = oracle.synth_corpus(SEED, 0, n, 384, f16_round=(dtype == "f16"))
    queries, near = oracle.synth_queries(SEED, n, 6, 384)
    assert_same(oracle, st, seen, queries, 10, variant=variant)
    assert_same(oracle, st, seen, queries[:2], 120, variant=variant)
    ids, _, _ = st.search_batch(queries, 1)
    for q in range(1, 6, 2):                          # planted neighbours are found
        assert ids[q, 0] == near[q]
    st.close()


def test_other_dims(sg, oracle):
    for dim, dtype in ((64, "f16"), (128, "f16"), (768, "f16"), (1024, "f16"), (32, "f32"), (96, "f32"), (768, "f32")):
        n = 3000
        corpus = oracle.synth_corpus(SEED, 0, n, dim)
        seen = oracle.round_f16(corpus) if dtype == "f16" else corpus
        st = sg.GpuVectorStore(dim, dtype)
        st.add_vectors(corpus)
        queries, _ = oracle.synth_queries(SEED, n, 3, dim)
        assert_same(oracle, st, seen, queries, 10)
        st.close()


def test_deleted_ids_never_returned(sg, oracle):
    """vamana.rs:1175-1177, :799-803; bug_regression_tests.rs:670 (soft delete)."""
    n = 5000
    corpus = oracle.synth_corpus(SEED, 0, n, 384)
    st = sg.GpuVectorStore(384, "f16")
    st.add_vectors(corpus)
    seen = oracle.round_f16(corpus)
    queries, _ = oracle.synth_queries(SEED, n, 4, 384)
    first, _ = oracle.brute_force_search(seen, queries[1], 5)
    deleted = [int(first[0]), int(first[2]), 0, 31, 32, 4999]
    for d in deleted:
        assert st.mark_deleted(d) is True
    assert st.mark_deleted(deleted[0]) is True         # already deleted: still true (id < len)
    assert st.mark_deleted(n) is False                 # unknown id -> false
    assert st.deleted_count() == len(deleted)
    assert st.is_deleted(31) and not st.is_deleted(30)
    assert_same(oracle, st, seen, queries, 10, deleted=deleted)
    ids, _, _ = st.search_batch(queries, 100)
    assert not set(deleted) & set(ids.ravel().tolist())
    st.clear_deleted()
    assert st.deleted_count() == 0
    assert_same(oracle, st, seen, queries, 10)
    st.close()


def test_k_larger_than_live_rows_and_empty_store(sg, oracle):
    """similarity.rs:108-115 (k > candidates), vamana.rs:766-768 (empty -> Ok([]))."""
    st = sg.GpuVectorStore(384, "f16")
    q, _ = oracle.synth_queries(SEED, 0, 2, 384)
    ids, dist, counts = st.search_batch(q, 10)
    assert counts.tolist() == [0, 0] and (ids == 0xFFFFFFFF).all()
    assert st.search(q[0], 5) == []
    corpus = oracle.synth_corpus(SEED, 0, 7, 384)
    st.add_vectors(corpus)
    st.mark_deleted(3)
    assert_same(oracle, st, oracle.round_f16(corpus), q, 10, deleted=[3])
    res = st.search(q[0], 10)
    assert len(res) == 6
    st.close()


def test_ties_break_by_id(sg, oracle):
    """vamana.rs:1184: equal distances are ordered by ascending id."""
    base = oracle.synth_corpus(SEED, 0, 40, 384)
    corpus = np.concatenate([base, base[:10], base[:10]])          # triplicates of rows 0..9
    st = sg.GpuVectorStore(384, "f32")
    st.add_vectors(corpus)
    res = st.search(base[3], 3)
    assert [r[0] for r in res] == [3, 43, 53]
    assert res[0][1] == res[1][1] == res[2][1]
    assert_same(oracle, st, corpus, base[:5], 60)
    st.close()


def test_vamana_unit_test_five_vectors(sg, oracle):
    """vamana.rs:1697-1711, zero-padded from R^4 to the minimum supported width (adding
    exact-zero products does not change any partial sum)."""
    vecs = np.zeros((5, 32), F)
    vecs[:, :4] = [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [0.5, 0.5, 0, 0]]
    q = np.zeros(32, F); q[:4] = [0.9, 0.1, 0, 0]
    st = sg.GpuVectorStore(32, "f32")
    st.build(vecs)
    res = st.search(q, 2)
    assert len(res) == 2 and res[0][0] == 0
    assert res[0][1] == float(F(-0.9))
    st.close()


def test_retrieval_test_vector_self_query(sg, oracle):
    """retrieval.rs:2430-2438 / :2482-2529: test_vector(i) retrieves its own id first."""
    from helpers import ref_test_vector
    corpus = np.stack([ref_test_vector(i, 384) for i in range(300)])
    st = sg.GpuVectorStore(384, "f32")
    st.add_vectors(corpus)
    for i in (0, 7, 299):
        res = st.search(corpus[i], 5)
        assert res[0][0] == i and abs(-res[0][1] - 1.0) < 1e-5
    assert_same(oracle, st, corpus, corpus[[0, 7, 299]], 10)
    st.close()


def test_distance_value_pins(sg):
    """distance_inline.rs:515-529 (dot [1..9].[9..1] = 165) through the scan, zero-padded."""
    row = np.zeros((1, 32), F); row[0, :9] = np.arange(9, 0, -1)
    q = np.zeros(32, F); q[:9] = np.arange(1, 10)
    st = sg.GpuVectorStore(32, "f32")
    st.add_vectors(row)
    assert st.search(q, 1) == [(0, -165.0)]
    st.close()


def test_dimension_mismatch_is_an_error(sg):
    st = sg.GpuVectorStore(384, "f16")
    with pytest.raises(sg.SgError):
        st.add_vectors(np.zeros((2, 100), F))
    with pytest.raises(sg.SgError):
        st.search_batch(np.zeros((1, 128), F), 3)
    with pytest.raises(sg.SgError):
        sg.GpuVectorStore(100, "f16")                   # unsupported width: error, not UB
    with pytest.raises(sg.SgError):
        st.search_batch(np.zeros((1, 384), F), 0)
    st.close()


def test_two_runs_byte_identical(sg, oracle):
    """Spirit of tests/recall_determinism.rs:66-99."""
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(SEED, 0, 50_000)
    q, _ = oracle.synth_queries(SEED, 50_000, 8, 384)
    a = st.search_batch(q, 50)
    b = st.search_batch(q, 50)
    for x, y in zip(a, b):
        assert x.tobytes() == y.tobytes()
    st.close()


def test_incremental_growth_keeps_rows_and_tombstones(sg, oracle):
    st = sg.GpuVectorStore(384, "f16")
    corpus = oracle.synth_corpus(SEED, 0, 6000, 384)
    st.add_vectors(corpus[:10])
    st.mark_deleted(4)
    for lo in range(10, 6000, 1500):                   # forces several reallocations
        st.add_vectors(corpus[lo:lo + 1500])
    assert len(st) == 6000 and st.is_deleted(4)
    q, _ = oracle.synth_queries(SEED, 6000, 3, 384)
    assert_same(oracle, st, oracle.round_f16(corpus), q, 10, deleted=[4])
    st.close()


def test_id_base_offsets_reported_ids(sg, oracle):
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(SEED, 1000, 2000)               # shard holding global rows 1000..2999
    st.set_id_base(1000)
    seen = oracle.synth_corpus(SEED, 1000, 2000, 384, f16_round=True)
    q, _ = oracle.synth_queries(SEED, 3000, 2, 384)
    ids, dist, _ = st.search_batch(q, 5)
    for i in range(2):
        rid, rdist = oracle.brute_force_search(seen, q[i], 5)
        assert np.array_equal(ids[i], rid + 1000) and np.array_equal(bits(dist[i]), bits(rdist))
    st.close()


def test_full_size_1m_properties(sg, oracle):
    """config[1] size (1M x 384 fp16): oracle comparison on 3 queries (CPU ~2 s) + size-independent
    properties: self-query returns own id with score == |row|^2 in the reference order, top-k is a
    prefix of top-(k+5), sharded == unsharded."""
    n = 1_000_000
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(SEED, 0, n)
    seen = oracle.synth_corpus(SEED, 0, n, 384, f16_round=True)
    q, near = oracle.synth_queries(SEED, n, 8, 384)
    assert_same(oracle, st, seen, q[:3], 10)
    rows = [5, 123_456, 999_999]
    ids, dist, _ = st.search_batch(seen[rows], 1)
    for j, r in enumerate(rows):
        assert ids[j, 0] == r
        assert dist[j, 0] == F(-oracle.dot(seen[r], seen[r]))
    a_ids, a_dist, _ = st.search_batch(q, 10)
    b_ids, b_dist, _ = st.search_batch(q, 15)
    assert np.array_equal(a_ids, b_ids[:, :10]) and np.array_equal(bits(a_dist), bits(b_dist[:, :10]))
    # every returned distance is the oracle's distance for that row
    for i in range(8):
        assert np.array_equal(bits(oracle.score_rows(seen, q[i], a_ids[i])), bits(a_dist[i]))
    # two half shards + merge == one store
    import torch
    halves = []
    for s0 in (0, n // 2):
        sh = sg.GpuVectorStore(384, "f16")
        sh.append_synthetic(SEED, s0, n // 2)
        sh.set_id_base(s0)
        halves.append(sh.search_batch(q, 10))
        sh.close()
    ids_l = torch.tensor(np.stack([halves[0][0], halves[1][0]], 1).astype(np.int64)).to(torch.int32).cuda()  # [b,2,k]
    dist_l = torch.tensor(np.stack([halves[0][1], halves[1][1]], 1)).cuda()
    out_i = torch.empty((8, 10), dtype=torch.int32, device="cuda")
    out_d = torch.empty((8, 10), dtype=torch.float32, device="cuda")
    out_c = torch.empty(8, dtype=torch.int32, device="cuda")
    sg.merge_topk_device(0, ids_l.data_ptr(), dist_l.data_ptr(), 8, 2, 10, out_i.data_ptr(), out_d.data_ptr(),
                         out_c.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(out_i.cpu().numpy().view(np.uint32), a_ids)
    assert np.array_equal(bits(out_d.cpu().numpy()), bits(a_dist))
    st.close()


def test_search_device_pointers_and_profiling(sg, oracle):
    import torch
    n = 200_000
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(SEED, 0, n)
    q, _ = oracle.synth_queries(SEED, n, 4, 384)
    dq = torch.tensor(q).cuda()
    ids = torch.empty((4, 10), dtype=torch.int32, device="cuda")
    dist = torch.empty((4, 10), dtype=torch.float32, device="cuda")
    cnt = torch.empty(4, dtype=torch.int32, device="cuda")
    st.set_profiling(True)
    st.set_tc_mode("exact")                            # 3 and 4 queries would take the tensor-core sweep by default
    before = sg.kernel_launch_count()
    st.search_device(dq.data_ptr(), 4, 10, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(),
                     torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert sg.kernel_launch_count() - before == 2      # one sweep (4 queries) + one merge
    ms, launches = st.last_scan_ms()
    assert launches == 1 and ms > 0
    h_ids, h_dist, _ = st.search_batch(q, 10)
    assert np.array_equal(ids.cpu().numpy().view(np.uint32), h_ids)
    assert np.array_equal(bits(dist.cpu().numpy()), bits(h_dist))
    st.close()


def test_k_1024_and_large_k_batches(sg, oracle):
    """Documented limit k <= 1024; k > 128 (any b) takes the exact sweep in groups of 4 queries."""
    n = 6000
    corpus = oracle.synth_corpus(SEED, 0, n, 384)
    seen = oracle.round_f16(corpus)
    st = sg.GpuVectorStore(384, "f16")
    st.add_vectors(corpus)
    q, _ = oracle.synth_queries(SEED, n, 7, 384)
    assert_same(oracle, st, seen, q[:1], 1024)
    assert_same(oracle, st, seen, q[:2], 500)
    assert_same(oracle, st, seen, q, 120)                 # the reference's k = 12 x max_results for top-10
    with pytest.raises(sg.SgError):
        st.search_batch(q[:1], 1025)
    st.close()


def test_concurrent_callers_serialise_correctly(sg, oracle):
    """The handle may be shared by many threads (spawn_blocking workers in the reference, recall.rs:513)."""
    import threading
    n = 50_000
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(SEED, 0, n)
    seen = oracle.synth_corpus(SEED, 0, n, 384, f16_round=True)
    q, _ = oracle.synth_queries(SEED, n, 16, 384)
    want = [oracle.brute_force_search(seen, q[i], 10) for i in range(16)]
    errors = []

    def worker(i):
        try:
            for _ in range(5):
                if i % 4 == 0:                             # some callers batch (tensor path), most do single queries
                    ids, dist, _ = st.search_batch(q[:8], 10)
                    for j in range(8):
                        assert np.array_equal(ids[j], want[j][0]) and np.array_equal(bits(dist[j]), bits(want[j][1]))
                else:
                    res = st.search(q[i], 10)
                    assert [r[0] for r in res] == want[i][0].tolist()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
    ts = [threading.Thread(target=worker, args=(i,)) for i in range(16)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors[:3]
    st.close()


def test_page_locked_queries_are_read_in_place(sg, oracle):
    """sg_store_search DMAs page-locked caller memory directly (no staging copy); same answers as pageable memory."""
    import torch
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(SEED, 0, 30000)
    corpus = oracle.synth_corpus(SEED, 0, 30000, 384, f16_round=True)
    rng = np.random.default_rng(11)
    q = rng.standard_normal((5, 384)).astype(F)
    pinned = torch.from_numpy(q.copy()).pin_memory()
    a = st.search_batch(pinned.numpy(), 10)
    b = st.search_batch(q, 10)
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x).view(np.uint32), np.asarray(y).view(np.uint32))
    assert_same(oracle, st, corpus, pinned.numpy(), 10)
    st.close()
