"""RetrievalEngine::search_ids post-processing on the device: chunk de-dup by MemoryId,
similarity = -distance, sort (similarity desc, MemoryId asc), truncate -- against the oracle's restatement of
retrieval.rs:927-961 fed with the oracle's own brute-force scan."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SEED = 23821
MAXU = np.iinfo(np.uint64).max


@pytest.mark.parametrize("b,limit", [(1, 10), (3, 25), (12, 10)])
def test_search_ids_matches_oracle(sg, oracle, b, limit):
    n = 20_000
    rng = np.random.default_rng(b * 100 + limit)
    corpus = oracle.synth_corpus(SEED, 0, n, 384)
    # chunked memories: ~3 vectors per memory, a few vectors unmapped, duplicated rows to force ties
    corpus[5000:5010] = corpus[42]
    n_mem = n // 3
    owner = rng.integers(0, n_mem, n)
    vec2mem = np.stack([np.full(n, 7, np.uint64), owner.astype(np.uint64) * np.uint64(1_000_003) % np.uint64(2**40)], 1)
    unmapped = rng.choice(n, 50, replace=False)
    vec2mem[unmapped] = MAXU
    st = sg.GpuVectorStore(384, "f16")
    st.add_vectors(corpus)
    st.set_memory_ids(vec2mem)
    seen = oracle.round_f16(corpus)
    queries, _ = oracle.synth_queries(SEED, n, b, 384)
    queries[0] = seen[42]
    mem, sim, cnt = st.search_ids(queries, limit)
    for q in range(b):
        rid, rdist = oracle.brute_force_search(seen, queries[q], 4 * limit)
        omem, osim = oracle.search_ids_post(rid, rdist, vec2mem, limit)
        c = int(cnt[q])
        assert c == len(omem)
        assert np.array_equal(mem[q, :c], omem), (q, mem[q, :c], omem)
        assert np.array_equal(sim[q, :c].view(np.uint32), osim.view(np.uint32))
    st.close()
