"""GPU parity for Product Quantization (reference src/vector_db/pq.rs): k-means training,
encode / decode, the ADC distance table and CompressedVectorStore::search must equal the CPU oracle bit for
bit (codes, table entries, distances, ids).  Everything goes through the C ABI (sg_pq_*)."""
import numpy as np
import pytest


pytestmark = pytest.mark.gpu
F = np.float32


def bits(a):
    return np.ascontiguousarray(a, F).view(np.uint32)


def test_training_encode_table_equal_the_oracle(sg, oracle):
    rng = np.random.default_rng(0)
    v = rng.random((2000, 384), dtype=F)
    ref = oracle.PQOracle(384, kmeans_iterations=5)
    orders = ref.shuffled_orders(2000, 7)
    ref.train(v, orders=orders)
    pq = sg.GpuProductQuantizer.train(v, orders, kmeans_iterations=5)
    assert (pq.num_subvectors, pq.num_centroids, pq.subvec_dim) == (48, 256, 8)
    assert np.array_equal(bits(pq.centroids), bits(ref.centroids))           # k-means, bit for bit
    x = rng.random((777, 384), dtype=F)
    codes = pq.encode(x)
    assert np.array_equal(codes, ref.encode(x))
    assert np.array_equal(bits(pq.decode(codes)), bits(ref.decode(codes)))
    q = rng.random(384, dtype=F)
    assert np.array_equal(bits(pq.build_distance_table(q)), bits(ref.table(q)))
    # pq.rs:497-521: reconstruction error of a training vector is small
    assert float(np.mean((v[0] - pq.decode(pq.encode(v[:1]))[0]) ** 2)) < 0.1
    pq.close()


def test_few_training_vectors_cap_the_centroids(sg, oracle):
    """pq.rs:118: num_centroids.min(n) (the reference's test_distance_table trains on 100 vectors)"""
    v = np.random.default_rng(1).random((100, 384), dtype=F)
    ref = oracle.PQOracle(384, kmeans_iterations=3)
    orders = ref.shuffled_orders(100, 1)
    ref.train(v, orders=orders)
    pq = sg.GpuProductQuantizer.train(v, orders, kmeans_iterations=3)
    assert pq.num_centroids == ref.k == 100
    assert np.array_equal(bits(pq.centroids), bits(ref.centroids))
    codes = pq.encode(v[1:2])
    t = pq.build_distance_table(v[0])
    tot = F(0)
    for s in range(48):
        tot = F(tot + t[s, codes[0, s]])
    assert float(tot) == ref.asymmetric_distance(v[0], codes[0])                # table == direct (pq.rs:534-551)
    pq.close()


@pytest.mark.parametrize("n", [1, 33, 5000, 30_000])
def test_compressed_store_search_bit_identical(sg, oracle, n):
    rng = np.random.default_rng(n)
    train = rng.random((1500, 384), dtype=F)
    ref = oracle.PQOracle(384, kmeans_iterations=2)
    ref.train(train, seed=3)
    pq = sg.GpuProductQuantizer.from_centroids(ref.centroids)
    v = rng.random((n, 384), dtype=F)
    ids = rng.permutation(np.arange(10, 10 + n)).astype(np.uint32)             # ids are labels, not row numbers
    half = n // 2
    pq.add(ids[:half], v[:half])
    pq.add(ids[half:], v[half:])
    ref.add(ids, v)
    assert len(pq) == n
    q = rng.random((5, 384), dtype=F)
    q[0] = v[0]
    for k in (1, 10, 100):
        gi, gd, gc = pq.search_batch(q, k)
        for i in range(5):
            ri, rd = ref.search(q[i], k)
            c = int(gc[i])
            assert c == len(ri) and np.array_equal(gi[i, :c], ri), (n, k, i, gi[i, :c], ri)
            assert np.array_equal(bits(gd[i, :c]), bits(rd))
            assert (gi[i, c:] == 0xFFFFFFFF).all()
    if n >= 5000:                                                              # pq.rs:554-577: the query is in the top 5
        assert int(ids[0]) in [i for i, _ in pq.search(v[0], 10)][:5]
    pq.close()


def test_small_shapes_and_ties(sg, oracle):
    """dim 32 (4 subvectors of 8, 16 centroids): many equal ADC distances -> ties are broken by id"""
    rng = np.random.default_rng(5)
    train = rng.standard_normal((300, 32)).astype(F)
    ref = oracle.PQOracle(32, num_centroids=16, kmeans_iterations=3)
    ref.train(train, seed=1)
    pq = sg.GpuProductQuantizer.from_centroids(ref.centroids)
    v = np.repeat(rng.standard_normal((50, 32)).astype(F), 20, axis=0)          # every vector 20 times
    ids = np.arange(1000, dtype=np.uint32)[::-1].copy()
    pq.add(ids, v)
    ref.add(ids, v)
    q = rng.standard_normal((3, 32)).astype(F)
    for k in (7, 64, 1000):
        gi, gd, gc = pq.search_batch(q, k)
        for i in range(3):
            ri, rd = ref.search(q[i], k)
            assert np.array_equal(gi[i, :gc[i]], ri) and np.array_equal(bits(gd[i, :gc[i]]), bits(rd)), (k, i)
    assert pq.search(q[0], 5) == [(int(a), float(b)) for a, b in zip(*ref.search(q[0], 5))]
    empty = sg.GpuProductQuantizer.from_centroids(ref.centroids)
    assert empty.search(q[0], 5) == []
    with pytest.raises(sg.SgError):
        sg.GpuProductQuantizer.from_centroids(np.zeros((3, 16, 8), F))         # 3 subvectors: not a multiple of 4
    empty.close()
    pq.close()


def test_member_list_training_path_equals_the_oracle(sg, oracle, monkeypatch):
    """large training sets (n >= SG_PQ_MEMBERS_MIN_N) take pq_update_members_kernel; forced here on a small set"""
    monkeypatch.setenv("SG_PQ_MEMBERS_MIN_N", "1")
    v = np.random.default_rng(9).random((1500, 384), dtype=F)
    ref = oracle.PQOracle(384, kmeans_iterations=4)
    orders = ref.shuffled_orders(1500, 2)
    ref.train(v, orders=orders)
    pq = sg.GpuProductQuantizer.train(v, orders, kmeans_iterations=4)
    assert np.array_equal(bits(pq.centroids), bits(ref.centroids))
    pq.close()
