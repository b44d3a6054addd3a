"""Pins the CPU oracle of the Vamana graph index (oracle/vamana_oracle.c) against the reference's own
unit tests (src/vector_db/vamana.rs:1685-1774) and against an independent pure-Python restatement of
greedy_search / robust_prune that uses heapq exactly where the reference uses BinaryHeap."""
import heapq

import numpy as np
import pytest

F = np.float32


def key(o, d, i):
    return (o.total_order_key(float(d)), int(i))


def greedy_py(o, V, query, k, entry):
    """vamana.rs:576-657 with two real heaps (candidates: min-heap; w: max-heap via negated keys)."""
    dist = lambda i: F(o.distance(query, V.vectors[i], V.metric, V.variant))
    visited = {entry}
    d0 = dist(entry)
    cand = [(key(o, d0, entry), entry, d0)]
    w = [((-key(o, d0, entry)[0], -entry), entry, d0)]
    while cand:
        _, cur, cd = heapq.heappop(cand)
        if w and cd > w[0][2]:
            break
        for v in V.nbr[cur, :V.deg[cur]]:
            v = int(v)
            if v in visited:
                continue
            visited.add(v)
            d = dist(v)
            if len(w) < k or d < w[0][2]:
                heapq.heappush(cand, (key(o, d, v), v, d))
                heapq.heappush(w, ((-key(o, d, v)[0], -v), v, d))
                if len(w) > k:
                    heapq.heappop(w)
    out = sorted(((key(o, d, i), i, d) for _, i, d in w))
    return [i for _, i, _ in out], [d for _, _, d in out]


def prune_py(o, V, node, cands):
    """vamana.rs:665-746."""
    cands = sorted(cands, key=lambda c: key(o, c[1], c[0]))
    off = F(1.0) if V.metric == 0 else F(0.0)
    kept, ne = [], []
    for cid, _ in cands:
        if cid == node:
            continue
        d_nc = F(o.distance(V.vectors[node], V.vectors[cid], V.metric, V.variant))
        add = True
        for e, d_ne in zip(kept, ne):
            d_ce = F(o.distance(V.vectors[cid], V.vectors[e], V.metric, V.variant))
            if F(F(V.alpha) * F(d_ce + off)) <= F(d_nc + off) and d_ce <= d_ne:
                add = False
                break
        if add:
            kept.append(cid)
            ne.append(d_nc)
            if len(kept) >= V.R:
                break
    return kept


def test_vamana_construction(oracle):
    """vamana.rs:1685-1711 test_vamana_construction: query [0.9,0.1,0,0] -> 2 results, closest id 0."""
    V = oracle.VamanaOracle(4, max_degree=3, search_list_size=10, alpha=1.2)
    V.build(np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [0.5, 0.5, 0, 0]], F))
    ids, dist = V.search([0.9, 0.1, 0, 0], 2)
    assert len(ids) == 2 and ids[0] == 0
    assert (V.deg <= 3).all()


@pytest.mark.parametrize("seed", range(5))
def test_estimate_recall(oracle, seed):
    """vamana.rs:1744-1774 test_estimate_recall: 10 vectors, R=4, L=20 -> recall@3 of 5 sampled rows >= 0.6
    (the reference samples rows at random; every start graph / sample must pass)."""
    vecs = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [.5, .5, 0, 0], [.5, 0, .5, 0],
                     [0, .5, .5, 0], [0, 0, .5, .5], [.25, .25, .25, .25], [.7, .3, 0, 0]], F)
    V = oracle.VamanaOracle(4, max_degree=4, search_list_size=20)
    V.build(vecs, seed=seed)
    rng = np.random.default_rng(seed)
    rec = []
    for qi in rng.permutation(10)[:5]:
        a, _ = V.search(vecs[qi], 3)
        e, _ = oracle.brute_force_search(vecs, vecs[qi], 3)
        rec.append(len(set(a) & set(e)) / 3)
    assert np.mean(rec) >= 0.6


@pytest.mark.parametrize("metric", [0, 1, 2])
def test_c_matches_python_restatement(oracle, metric):
    x = oracle.clustered_corpus(3, 400, 32)
    if metric:
        x = (x * np.random.default_rng(1).uniform(0.5, 2, (400, 1))).astype(F)
    V = oracle.VamanaOracle(32, max_degree=8, search_list_size=20, metric=metric)
    V.build(x, seed=9)
    assert (V.deg <= 8).all() and (V.deg > 0).all()
    rng = np.random.default_rng(2)
    for t in range(12):
        q = (x[rng.integers(0, 400)] + F(0.05) * rng.standard_normal(32)).astype(F)
        for k in (1, 5, 30):
            ids, dist = V.greedy_search(q, k)
            pids, pdist = greedy_py(oracle, V, q, k, V.medoid)
            assert list(ids) == pids and [float(d) for d in dist] == [float(d) for d in pdist]
        node = int(rng.integers(0, 400))
        ids, dist = V.greedy_search(x[node], 20)
        assert list(V.robust_prune(node, ids, dist)) == prune_py(oracle, V, node, list(zip(ids.tolist(), dist.tolist())))


def test_build_is_a_fixed_point_of_its_own_rule(oracle):
    """every list the build leaves is <= R long, has no self loop and no duplicate"""
    x = oracle.clustered_corpus(4, 600, 32)
    V = oracle.VamanaOracle(32, max_degree=12, search_list_size=30)
    assert V.build(x) in (1, 2)                       # vamana.rs:276: at most two passes
    for i in range(600):
        row = V.nbr[i, :V.deg[i]]
        assert i not in row and len(set(row.tolist())) == len(row) and V.deg[i] <= 12


def test_search_skips_deleted_and_widens_the_beam(oracle):
    """vamana.rs:785-805: search_k = k + min(deleted, 2k); deleted ids never come back."""
    x = oracle.clustered_corpus(5, 500, 32)
    V = oracle.VamanaOracle(32, max_degree=12, search_list_size=30)
    V.build(x)
    q = x[7]
    before, _ = V.search(q, 5)
    assert before[0] == 7
    V.mark_deleted(7)
    V.mark_deleted(int(before[1]))
    after, dist = V.search(q, 5)
    assert 7 not in after and before[1] not in after and len(after) == 5
    wide, wdist = V.greedy_search(q, 5 + 2)           # the widened beam, then filtered
    keep = [i for i in wide if i not in V.deleted][:5]
    assert list(after) == keep


def test_add_vector_incremental(oracle):
    """vamana.rs:1713-1742 shape: build 3, add 5 -> ids 3..7, all searchable, degrees bounded."""
    V = oracle.VamanaOracle(4, max_degree=3, search_list_size=10)
    V.build(np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]], F))
    for i in range(5):
        assert V.add_vector([0.1 * i, 0.1, 0.1, 0.1]) == 3 + i
    assert len(V) == 8 and (V.deg <= 3).all()
    for k in (1, 3, 8):                               # the grown graph searches like the restatement says
        ids, _ = V.search([0.4, 0.1, 0.1, 0.1], k)
        assert list(ids) == greedy_py(oracle, V, np.array([0.4, 0.1, 0.1, 0.1], F), k, V.medoid)[0]
    E = oracle.VamanaOracle(4, max_degree=3)          # first vector of an empty index: medoid 0, no edges
    assert E.add_vector([1, 0, 0, 0]) == 0 and E.deg[0] == 0 and E.medoid == 0
    assert E.add_vector([0, 1, 0, 0]) == 1 and list(E.nbr[1, :E.deg[1]]) == [0] and list(E.nbr[0, :E.deg[0]]) == [1]


def test_find_medoid(oracle):
    x = oracle.clustered_corpus(6, 300, 16)
    cen = np.zeros(16, F)
    for r in x:
        cen = (cen + r).astype(F)
    cen = (cen / F(300)).astype(F)
    d = [oracle.distance(r, cen, 0, 0) for r in x]
    assert oracle.find_medoid(x) == int(np.argmin(d))


# ---- the device kernel's formulation of greedy_search, modelled step by step -------------------------
fast_steps, slow_steps = [0], [0]


def greedy_single_array_py(o, V, query, K, entry, slack=32, batched=True):
    """shodh_memory_b200/csrc/vamana.cu::ann_search_kernel in pure Python: ONE sorted array P that is both
    heaps of the reference (w = its first K entries, candidates = its unexpanded entries); entries beyond
    the K-th survive only while they tie the K-th distance.  Must give the reference's result and traversal."""
    dist = lambda i: F(o.distance(query, V.vectors[i], V.metric, V.variant))
    tk = lambda d, i: (o.total_order_key(float(d)), int(i))
    d0 = dist(entry)
    P = [[tk(d0, entry), d0, entry, False]]           # key, dist, id, expanded
    visited = {entry}
    hops, evals = 0, 1
    while True:
        cur = next((e for e in P if not e[3]), None)
        if cur is None:
            break
        wl = min(len(P), K)
        if cur[1] > P[wl - 1][1]:
            break
        cur[3] = True
        hops += 1
        nbrs = [int(v) for v in V.nbr[cur[2], :V.deg[cur[2]]]]
        for c0 in range(0, len(nbrs), 32):            # one neighbour per lane
            lanes = []
            for v in nbrs[c0:c0 + 32]:
                if v not in visited:
                    visited.add(v)
                    lanes.append((v, dist(v)))
                    evals += 1
            if len(P) >= K:                            # pre-filter against the K-th distance as it stands
                lanes = [(v, d) for v, d in lanes if d < P[K - 1][1]]
            # batched merge (the kernel's fast path): with no float-equal distances next to a new entry or across
            # the K-th boundary, the result of the sequential rule is the first K of sorted(P + candidates)
            if batched and len(lanes) > 2 and len(P) <= K and not any(np.isnan(d) for _, d in lanes) \
                    and not np.isnan(P[-1][1]):
                M = sorted(P + [[tk(d, v), d, v, 2] for v, d in lanes], key=lambda e: e[0])[:K + 1]
                tie = any(M[i][1] == M[i + 1][1] and (M[i][3] == 2 or M[i + 1][3] == 2 or i == K - 1)
                          for i in range(len(M) - 1))
                if not tie:
                    P = [[e[0], e[1], e[2], False if e[3] == 2 else e[3]] for e in M[:K]]
                    fast_steps[0] += 1
                    continue
                slow_steps[0] += 1
            for v, d in lanes:                         # lane order
                if len(P) >= K and not (d < P[K - 1][1]):
                    continue
                e = [tk(d, v), d, v, False]
                pos = sum(1 for x in P if x[0] < e[0])
                P.insert(pos, e)
                if len(P) > K:
                    wd = P[K - 1][1]
                    cut = next((i for i in range(K, len(P)) if P[i][1] > wd), None)
                    if cut is not None:
                        del P[cut:]
                    assert len(P) < K + slack
    w = P[:min(len(P), K)]
    return [e[2] for e in w], [e[1] for e in w], (hops, evals)


@pytest.mark.parametrize("dups", [False, True])
def test_single_array_formulation_equals_two_heaps(oracle, dups):
    x = oracle.clustered_corpus(17, 300, 16)
    if dups:                                           # exact distance ties: every row of the first 100 three times
        x = np.vstack([x, x[:100], x[:100]]).astype(F)
    V = oracle.VamanaOracle(16, max_degree=6, search_list_size=16)
    V.build(x, seed=4)
    rng = np.random.default_rng(3)
    for t in range(30):
        q = x[rng.integers(0, len(x))] if dups else (x[rng.integers(0, len(x))] + F(0.1) * rng.standard_normal(16)).astype(F)
        for k in (1, 2, 4, 9, 33):
            ids, dist, st = V.greedy_search(q, k, stats=True)
            for batched in (False, True):
                sids, sdist, sst = greedy_single_array_py(oracle, V, q, k, V.medoid, batched=batched)
                assert list(ids) == sids and [float(d) for d in dist] == [float(d) for d in sdist], (t, k, batched)
                assert (int(st[0]), int(st[1])) == sst, (t, k, batched, st, sst)
    assert fast_steps[0] > 0 and (slow_steps[0] > 0 or not dups)      # both paths were exercised
