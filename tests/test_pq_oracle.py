"""Pins the CPU oracle of Product Quantization (oracle/pq_oracle.c) against the reference's own unit tests
(src/vector_db/pq.rs:484-578) and an independent numpy-f32 restatement."""
import numpy as np

F = np.float32


def sq_l2_py(a, b):
    s = F(0)
    for x, y in zip(a, b):
        d = F(x - y)
        s = F(s + F(d * d))
    return s


def test_pq_encode_decode(oracle):
    """pq.rs:497-521 test_pq_encode_decode: 1000 x 384 uniform vectors, 48 codes, MSE < 0.1"""
    v = np.random.default_rng(0).random((1000, 384), dtype=F)
    pq = oracle.PQOracle(384).train(v, seed=1)
    codes = pq.encode(v[:1])
    dec = pq.decode(codes)
    assert codes.shape == (1, 48) and dec.shape == (1, 384)
    assert float(np.mean((v[0] - dec[0]) ** 2)) < 0.1


def test_compression_ratio(oracle):
    """pq.rs:524-531: 1536 bytes -> 48 bytes = 32x"""
    pq = oracle.PQOracle(384)
    assert pq.dim * 4 == 1536 and pq.m == 48 and abs(pq.dim * 4 / pq.m - 32.0) < 0.01


def test_distance_table(oracle):
    """pq.rs:534-551: table distance == direct asymmetric distance (n = 100 < 256 -> k = 100 centroids)"""
    v = np.random.default_rng(1).random((100, 384), dtype=F)
    pq = oracle.PQOracle(384, kmeans_iterations=5).train(v, seed=2)
    assert pq.k == 100
    codes = pq.encode(v[1:2])[0]
    direct = pq.asymmetric_distance(v[0], codes)
    via_table = pq.distance_with_table(pq.table(v[0]), codes)
    assert abs(direct - via_table) < 1e-6 and direct == via_table      # same additions in the same order


def test_compressed_store_search(oracle):
    """pq.rs:554-577: the query vector itself is within the top 5 of 10 results"""
    v = np.random.default_rng(2).random((1000, 384), dtype=F)
    pq = oracle.PQOracle(384, kmeans_iterations=4).train(v, seed=3)
    pq.add(np.arange(1000), v)
    ids, dist = pq.search(v[0], 10)
    assert len(ids) == 10 and 0 in ids[:5]
    assert all(dist[i] <= dist[i + 1] for i in range(9))


def test_c_matches_python_restatement(oracle):
    rng = np.random.default_rng(3)
    v = rng.standard_normal((300, 32)).astype(F)
    pq = oracle.PQOracle(32, subvec_dim=8, num_centroids=16, kmeans_iterations=3)
    orders = pq.shuffled_orders(300, 5)
    pq.train(v, orders=orders)
    # k-means restated: init = first k of the shuffled order, assign (strict <), mean in vector order
    for s in range(4):
        sub = v[:, s * 8:(s + 1) * 8]
        cent = sub[orders[s][:16]].copy()
        for _ in range(3):
            assign = [int(np.argmin([sq_l2_py(x, c) for c in cent])) for x in sub]
            acc = np.zeros((16, 8), F); cnt = np.zeros(16, np.int64)
            for x, a in zip(sub, assign):
                cnt[a] += 1
                acc[a] = (acc[a] + x).astype(F)
            for c in range(16):
                if cnt[c]:
                    cent[c] = (acc[c] / F(cnt[c])).astype(F)
        assert np.array_equal(cent.view(np.uint32), pq.centroids[s].view(np.uint32)), s
    q = rng.standard_normal(32).astype(F)
    t = pq.table(q)
    for s in range(4):
        for c in range(16):
            assert t[s, c] == sq_l2_py(q[s * 8:(s + 1) * 8], pq.centroids[s, c])
    codes = pq.encode(v[:20])
    for i in range(20):
        for s in range(4):
            d = [sq_l2_py(v[i, s * 8:(s + 1) * 8], c) for c in pq.centroids[s]]
            assert codes[i, s] == int(np.argmin(d))                    # argmin = first strict minimum
    pq.add(np.arange(300)[::-1], v)                                    # ids need not be row numbers
    ids, dist = pq.search(q, 7)
    all_d = []
    for i in range(300):
        tot = F(0)
        for s in range(4):
            tot = F(tot + t[s, pq.codes[i, s]])
        all_d.append((oracle.total_order_key(float(tot)), int(pq.ids[i]), tot))
    all_d.sort()
    assert [a[1] for a in all_d[:7]] == list(ids) and [a[2] for a in all_d[:7]] == list(dist)
