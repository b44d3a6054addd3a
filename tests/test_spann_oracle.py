"""Pins the CPU oracle of the SPANN index (oracle/spann_oracle.c) against the reference's own unit tests
(src/vector_db/spann.rs:1100-1238) and an independent numpy restatement of the search."""
import numpy as np
import pytest

F = np.float32


def unit_rows(seed, n, dim):
    """generate_random_vectors (:1105-1119): uniform [0,1) rows, L2-normalised"""
    v = np.random.default_rng(seed).random((n, dim), dtype=F)
    return np.ascontiguousarray(v / np.linalg.norm(v, axis=1, keepdims=True), F)


def test_partition_count(oracle):
    """spann.rs:1186-1192"""
    s = oracle.SpannOracle()
    assert (s.compute_partitions(100), s.compute_partitions(10000), s.compute_partitions(1000000)) == (10, 100, 1000)


def test_spann_build_and_search(oracle):
    """spann.rs:1122-1150: 1000 x 384, 20 probes -> exactly 10 results, the query vector within the top 3"""
    v = unit_rows(0, 1000, 384)
    s = oracle.SpannOracle(384, num_probes=20, kmeans_iterations=8, pq_iterations=4).build(v, seed=1)
    assert s.centroids.shape[0] == 32 and len(s) == 1000
    ids, dist = s.search(v[0], 10)
    assert len(ids) == 10 and 0 in ids[:3]
    assert all(dist[i] <= dist[i + 1] for i in range(9))


def test_build_rejects_use_pq_false_and_bad_dimensions(oracle):
    """spann.rs:1195-1237"""
    v = unit_rows(1, 100, 384)
    with pytest.raises(ValueError, match="use_pq=true"):
        oracle.SpannOracle(384, use_pq=False).build(v)
    s = oracle.SpannOracle(384, kmeans_iterations=3, pq_iterations=2).build(v)
    with pytest.raises(ValueError, match="dimension"):
        s.search(np.full(128, 0.1, F), 5)
    with pytest.raises(ValueError):
        oracle.SpannOracle(384).build(np.zeros((0, 384), F))


@pytest.mark.parametrize("metric", [0, 1])
def test_c_matches_numpy_restatement(oracle, metric):
    rng = np.random.default_rng(4)
    v = unit_rows(2, 600, 32)
    s = oracle.SpannOracle(32, num_probes=4, kmeans_iterations=6, metric=metric, pq_iterations=3)
    s.pq_cfg = None
    s.build(v, seed=3)
    # k-means restated: init = first k of the order, assign by compute_distance (strict <), mean in vector order
    def dist(a, b):
        acc = F(0)
        if metric == 1:
            for x, y in zip(a, b):
                d = F(x - y); acc = F(acc + F(d * d))
            return acc
        for x, y in zip(a, b):
            acc = F(acc + F(x * y))
        return F(F(1) - acc)
    k = s.centroids.shape[0]
    cent = v[s.order[:k]].copy()
    assign = np.zeros(600, np.int64)
    for _ in range(6):
        new = np.array([int(np.argmin([dist(x, c) for c in cent])) for x in v])
        changed = int((new != assign).sum())
        assign = new
        acc = np.zeros((k, 32), F); cnt = np.zeros(k, np.int64)
        for x, a in zip(v, assign):
            cnt[a] += 1
            acc[a] = (acc[a] + x).astype(F)
        for c in range(k):
            if cnt[c]:
                cent[c] = (acc[c] / F(cnt[c])).astype(F)
        if changed == 0:
            break
    assert np.array_equal(cent.view(np.uint32), s.centroids.view(np.uint32))
    # search restated
    q = (v[7] + F(0.01) * rng.standard_normal(32)).astype(F)
    ids, dd, probes = s.search(q, 9, with_probes=True)
    pr = sorted(range(k), key=lambda p: (oracle.total_order_key(float(dist(q, s.centroids[p]))), p))[:4]
    assert list(probes) == pr
    table = s.pq.table(q)
    cands = []
    for p in pr:
        for i, code in zip(s.part_ids[p], s.part_codes[p]):
            tot = F(0)
            for m in range(s.pq.m):
                tot = F(tot + table[m, code[m]])
            cands.append((oracle.total_order_key(float(tot)), int(i), tot))
    cands.sort()
    assert [c[1] for c in cands[:9]] == list(ids) and [c[2] for c in cands[:9]] == list(dd)
    # insert: lands in the nearest partition, found again by a search
    new = (v[11] * F(1.0)).copy()
    p = s.insert(5000, new)
    assert p == int(np.argmin([dist(new, c) for c in s.centroids])) and len(s) == 601
    ids2, _ = s.search(new, 3)
    assert 5000 in ids2
