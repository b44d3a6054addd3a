"""Row storage grows IN PLACE (CUDA virtual memory management): no second copy of the corpus, base address fixed.

The reference's VamanaIndex grows its mmap'd vector file in place (vamana.rs:853-974 add_vector -> storage append);
the GPU store must not need 2x the corpus in HBM to take one more row.  Results are checked against the oracle
after every growth step, on both search paths.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F = np.float32
SEED = 7711


def bits(a):
    return np.ascontiguousarray(a, F).view(np.uint32)


def test_incremental_growth_matches_oracle(sg, oracle):
    dim, k = 384, 10
    st = sg.GpuVectorStore(dim, "f16")
    sizes = [1, 999, 3000, 12000, 40000, 100000]          # cumulative row counts after each append
    done = 0
    rng = np.random.default_rng(5)
    queries = rng.standard_normal((8, dim)).astype(F)
    infos = []
    for n in sizes:
        st.append_synthetic(SEED, done, n - done)
        done = n
        infos.append(st.memory_info())
        corpus = oracle.synth_corpus(SEED, 0, n, dim, f16_round=True)
        for mode in ("exact", "auto"):
            st.set_tc_mode(mode)
            ids, dist, counts = st.search_batch(queries, k)
            for q in range(queries.shape[0]):
                rid, rdist = oracle.brute_force_search(corpus, queries[q], k)
                c = int(counts[q])
                assert c == len(rid)
                assert np.array_equal(ids[q, :c], rid), (n, mode, q)
                assert np.array_equal(bits(dist[q, :c]), bits(rdist)), (n, mode, q)
    assert all(i["capacity_rows"] >= n for i, n in zip(infos, sizes))
    if infos[0]["in_place"]:
        assert all(i["in_place"] for i in infos)
        assert infos[-1]["chunks"] > 1                                   # grew by mapping, not by re-allocating
        assert infos[-1]["bytes"] < 2 * sizes[-1] * dim * 2 + (8 << 20)  # never anything like 2x the corpus
    st.close()


def test_growth_keeps_rows_and_deletions(sg):
    dim = 128
    st = sg.GpuVectorStore(dim, "f32")
    rng = np.random.default_rng(9)
    a = rng.standard_normal((5000, dim)).astype(F)
    st.add_vectors(a)
    st.mark_deleted(17)
    st.mark_deleted(4999)
    b = rng.standard_normal((200000, dim)).astype(F)
    st.add_vectors(b)                                                    # forces several growth steps
    assert len(st) == 205000
    assert np.array_equal(bits(st.get_vectors(0, 5000)), bits(a))
    assert np.array_equal(bits(st.get_vectors(5000, 200000)), bits(b))
    assert st.is_deleted(17) and st.is_deleted(4999) and not st.is_deleted(5000)
    assert st.deleted_count() == 2
    st.close()


def test_many_small_stores_stay_small(sg):
    """one index per user (reference: MultiUserMemoryManager) -- a small store must not pin tens of MB"""
    stores = [sg.GpuVectorStore(384, "f16") for _ in range(64)]
    for i, st in enumerate(stores):
        st.append_synthetic(SEED + i, 0, 100)
    assert all(st.memory_info()["bytes"] <= (4 << 20) for st in stores)
    for st in stores:
        st.close()


@pytest.mark.parametrize("env,in_place", [({"SG_STORE_VMM": "0"}, False), ({"SG_STORE_VA_MB": "8"}, True)])
def test_growth_modes_in_a_child_process(sg, oracle, env, in_place):
    """SG_STORE_VMM=0 forces the allocate-and-copy path; SG_STORE_VA_MB=8 makes the store outgrow its address range
    several times (chunks re-mapped into a larger range, nothing copied).  Both are read once per process -> child."""
    code = r"""
import numpy as np, sys
sys.path.insert(0, %r)
import shodh_memory_b200 as sg
from oracle import oracle
st = sg.GpuVectorStore(384, "f16")
done = 0
q = np.random.default_rng(5).standard_normal((4, 384)).astype(np.float32)
for n in (10, 5000, 60000, 150000):
    st.append_synthetic(7711, done, n - done); done = n
    ids, dist, counts = st.search_batch(q, 10)          # searches between growth steps keep working
assert st.memory_info()["in_place"] == %r, st.memory_info()
corpus = oracle.synth_corpus(7711, 0, done, 384, f16_round=True)
assert np.array_equal(st.get_vectors(0, done).view(np.uint32), corpus.view(np.uint32))
for mode in ("exact", "auto"):
    st.set_tc_mode(mode)
    ids, dist, counts = st.search_batch(q, 10)
    for i in range(4):
        rid, rdist = oracle.brute_force_search(corpus, q[i], 10)
        assert np.array_equal(ids[i], rid) and np.array_equal(dist[i].view(np.uint32), np.asarray(rdist, np.float32).view(np.uint32))
print("ok")
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), in_place)
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
