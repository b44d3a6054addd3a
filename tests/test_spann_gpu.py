"""GPU parity for the SPANN index (reference src/vector_db/spann.rs): build (both k-means),
probe selection, ADC search and insert must equal the CPU oracle bit for bit.  Everything goes through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F = np.float32


def bits(a):
    return np.ascontiguousarray(a, F).view(np.uint32)


def unit_rows(seed, n, dim):
    v = np.random.default_rng(seed).random((n, dim), dtype=F)
    return np.ascontiguousarray(v / np.linalg.norm(v, axis=1, keepdims=True), F)


@pytest.mark.parametrize("metric,mid", [("neg_dot", 0), ("euclidean", 1)])
def test_build_and_search_equal_the_oracle(sg, oracle, metric, mid):
    v = unit_rows(0, 3000, 384)
    ref = oracle.SpannOracle(384, num_probes=8, kmeans_iterations=6, metric=mid, pq_iterations=3).build(v, seed=2)
    idx = sg.GpuSpannIndex.build(v, ref.order, ref.pq_orders, num_probes=8, kmeans_iterations=6, metric=metric, pq_iterations=3)
    inf = idx.info()
    assert inf["num_partitions"] == ref.centroids.shape[0] == 55 and inf["len"] == 3000
    cent, pqc, off, ids, codes = idx.arrays(ref.pq.centroids.shape)
    assert np.array_equal(bits(cent), bits(ref.centroids))                   # partition k-means, bit for bit
    assert np.array_equal(bits(pqc), bits(ref.pq.centroids))                 # PQ k-means, bit for bit
    roff, rids, rcodes = ref.csr()
    assert np.array_equal(off, roff) and np.array_equal(ids, rids) and np.array_equal(codes, rcodes)
    rng = np.random.default_rng(5)
    q = np.ascontiguousarray(v[rng.integers(0, 3000, 12)] + F(0.01) * rng.standard_normal((12, 384)).astype(F), F)
    for k in (1, 10, 100):
        gi, gd, gc = idx.search_batch(q, k)
        for i in range(12):
            ri, rd = ref.search(q[i], k)
            c = int(gc[i])
            assert c == len(ri) and np.array_equal(gi[i, :c], ri), (metric, k, i)
            assert np.array_equal(bits(gd[i, :c]), bits(rd))
            assert (gi[i, c:] == 0xFFFFFFFF).all()
    idx.close()


def test_reference_unit_tests(sg, oracle):
    """spann.rs:1122-1150 (query within the top 3 of 10 results with 20 probes) and :1218-1237 (dimension mismatch)"""
    v = unit_rows(1, 1000, 384)
    ref = oracle.SpannOracle(384, num_probes=20, kmeans_iterations=8, pq_iterations=4).build(v, seed=1)
    idx = sg.GpuSpannIndex.build(v, ref.order, ref.pq_orders, num_probes=20, kmeans_iterations=8, pq_iterations=4)
    r = idx.search(v[0], 10)
    assert len(r) == 10 and 0 in [i for i, _ in r][:3]
    assert r == [(int(a), float(b)) for a, b in zip(*ref.search(v[0], 10))]
    with pytest.raises(sg.SgError):
        idx.search_batch(np.full((1, 128), 0.1, F), 5)
    idx.close()


def test_insert_and_from_arrays(sg, oracle):
    v = unit_rows(2, 1500, 384)
    ref = oracle.SpannOracle(384, num_probes=6, kmeans_iterations=4, pq_iterations=2).build(v, seed=4)
    off, ids, codes = ref.csr()
    idx = sg.GpuSpannIndex.from_arrays(ref.centroids, ref.pq.centroids, off, ids, codes, num_probes=6)
    extra = unit_rows(3, 40, 384)
    for j, x in enumerate(extra):
        idx.insert(10_000 + j, x)
        ref.insert(10_000 + j, x)
    assert len(idx) == 1540
    _, _, goff, gids, gcodes = idx.arrays(ref.pq.centroids.shape)
    roff, rids, rcodes = ref.csr()
    assert np.array_equal(goff, roff) and np.array_equal(gids, rids) and np.array_equal(gcodes, rcodes)
    gi, gd, gc = idx.search_batch(extra[:8], 5)
    for i in range(8):
        ri, rd = ref.search(extra[i], 5)
        assert np.array_equal(gi[i, :gc[i]], ri) and np.array_equal(bits(gd[i, :gc[i]]), bits(rd))
        assert 10_000 + i in gi[i]
    idx.set_probes(2)
    ref.num_probes = 2
    gi, gd, gc = idx.search_batch(extra[:4], 5)
    for i in range(4):
        ri, rd = ref.search(extra[i], 5)
        assert np.array_equal(gi[i, :gc[i]], ri)
    idx.close()
