"""BASELINE-size parity: 10M x 384 fp16 rows = 7.68 GB, i.e. byte offsets beyond 2^32 (row 5,592,406 onwards) --
exactly where an addressing bug would live.  (i) the tensor-core path must equal the exact sweep byte for byte on
256 queries (and at the reference's k = 120 on 32); (ii) both must equal the CPU oracle on queries whose answers
sit past the 2^32-byte mark, with soft-deleted rows among the planted neighbours and in the last tile.
The oracle never holds the corpus: it regenerates it chunk by chunk (oracle.brute_force_search_synth)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F = np.float32
SEED = 23821
N = 10_000_000
FOUR_GB_ROW = (1 << 32) // 768 + 1


@pytest.fixture(scope="module")
def big_store(sg):
    st = sg.GpuVectorStore(384, "f16")
    st.reserve(N)
    st.append_synthetic(SEED, 0, N)
    yield st
    st.close()


def test_10m_tensor_path_equals_exact_sweep_and_oracle(sg, oracle, big_store):
    st = big_store
    b = 256
    queries, near = oracle.synth_queries(SEED, N, b, 384)        # odd queries sit next to row near[q], spread over [0, N)
    assert (near[1::2] > FOUR_GB_ROW).sum() > 20                 # plenty of answers beyond byte offset 2^32
    # self-queries of rows in the upper 4M rows, one of them in the very last tile
    for i, row in enumerate((6_000_000, 7_777_777, 9_000_001, N - 3)):
        queries[2 * i] = st.get_vectors(row, 1)[0]
    deleted = [int(near[1]), int(near[3]), N - 1, N - 2, N - 200, 5_592_405, 5_592_406]
    for d in deleted:
        st.mark_deleted(d)
    before = st.tc_stats()
    tc = st.search_batch(queries, 10)
    after = st.tc_stats()
    assert after[0] - before[0] == b and after[1] - before[1] <= 4
    st.set_tc_mode("exact")
    ex = st.search_batch(queries, 10)
    st.set_tc_mode("auto")
    for x, y in zip(tc, ex):
        assert x.tobytes() == y.tobytes()
    ids = tc[0]
    assert ids[0, 0] == 6_000_000 and ids[2, 0] == 7_777_777 and ids[4, 0] == 9_000_001 and ids[6, 0] == N - 3
    for q in range(5, b, 2):                                     # planted neighbours (1 and 3 were deleted)
        assert ids[q, 0] == near[q]
    assert ids[1, 0] != near[1] and ids[3, 0] != near[3]
    # the oracle on 8 of them (4 self-queries high up, 2 with their neighbour deleted, 2 planted beyond 4 GB)
    high = [q for q in range(5, b, 2) if near[q] > FOUR_GB_ROW][:2]
    sel = [0, 2, 4, 6, 1, 3] + high
    rid, rdist, rcnt = oracle.brute_force_search_synth(SEED, N, queries[sel], 10, deleted=deleted)
    for j, q in enumerate(sel):
        assert rcnt[j] == 10
        assert np.array_equal(ids[q], rid[j]), (q, ids[q], rid[j])
        assert np.array_equal(tc[1][q].view(np.uint32), rdist[j].view(np.uint32)), (q, tc[1][q], rdist[j])
    st.clear_deleted()


def test_10m_reference_call_shape_k120(sg, oracle, big_store):
    """k = 120 (what search_ids asks for a top-10 recall): tensor-core path == exact sweep at 10M rows."""
    st = big_store
    b, k = 32, 120
    queries = oracle.synth_queries(SEED, N, 256 + b, 384)[0][256:]          # fresh queries of the same generator
    before = st.tc_stats()
    tc = st.search_batch(queries, k)
    after = st.tc_stats()
    assert after[0] - before[0] == b and after[1] - before[1] <= 2
    st.set_tc_mode("exact")
    ex = st.search_batch(queries, k)
    st.set_tc_mode("auto")
    for x, y in zip(tc, ex):
        assert x.tobytes() == y.tobytes()
    rid, rdist, _ = oracle.brute_force_search_synth(SEED, N, queries[:2], k)
    assert np.array_equal(tc[0][:2], rid) and np.array_equal(tc[1][:2].view(np.uint32), rdist.view(np.uint32))


def test_10m_batch1_top100_equals_oracle(sg, oracle, big_store):
    """configs[2]: batch-1 top-100 (exact sweep, single query) at full size."""
    st = big_store
    queries, near = oracle.synth_queries(SEED, N, 400, 384)
    q = queries[399:400]                                                     # odd index: planted next to row near[399]
    ids, dist, cnt = st.search_batch(q, 100)
    rid, rdist, _ = oracle.brute_force_search_synth(SEED, N, q, 100)
    assert cnt[0] == 100 and ids[0, 0] == near[399]
    assert np.array_equal(ids, rid) and np.array_equal(dist.view(np.uint32), rdist.view(np.uint32))
