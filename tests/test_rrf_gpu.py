"""GPU parity for RRFusion::fuse (reference src/memory/hybrid_search.rs:537-594): the
hand-off from the retrieval legs to the fused candidate list.  f32 divisions and additions in the
reference's order -> ids AND score bits must equal the oracle's exactly."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def same(got, want):
    gi, gs = got
    wi, ws = want
    assert gi.shape == wi.shape and np.array_equal(gi, wi)
    assert np.array_equal(gs.view(np.uint32), ws.view(np.uint32))


def pool(rng, n):
    return rng.integers(0, 2**64 - 1, (n, 2), dtype=np.uint64)


def test_reference_unit_cases(sg, oracle):
    rng = np.random.default_rng(1)
    a, b, c = pool(rng, 3)
    cases = [
        (60.0, [0.5, 0.5], [np.stack([a, b, c]), np.stack([b, a, c])]),      # hybrid_search.rs:969-1011
        (60.0, [0.5, 0.5], [a[None], b[None]]),                              # :1013-1029
        (60.0, [0.5, 0.5], [np.full((1, 2), 2**64 - 1, np.uint64), np.zeros((1, 2), np.uint64)]),   # :1031-1052
        (60.0, [0.8, 0.2], [np.stack([a, b]), np.stack([b, a])]),            # :1172-1196
        (1.0, [0.5, 0.5], [np.stack([a, b, c]), np.stack([c, b, a])]),       # :1198-1240
        (45.0, [0.35, 0.40], [np.stack([a, b, c]), np.stack([c, a])]),       # HybridSearchConfig defaults :67-78
    ]
    for k, w, lists in cases:
        same(sg.RRFusion(k, w).fuse(lists), oracle.rrf_fuse(k, w, lists))
    ids, sc = sg.RRFusion(60.0, [0.5, 0.5]).fuse(cases[2][2])
    assert np.array_equal(ids[0], [0, 0]) and sc[0] == sc[1]                 # tie -> lower MemoryId first


@pytest.mark.parametrize("sizes,overlap", [((100, 100), 0.5), ((240, 120), 0.2), ((1, 0), 0.0), ((0, 0), 0.0),
                                           ((700, 700, 648), 0.7), ((1024, 1024), 1.0), ((50, 50, 50, 50, 50), 0.4)])
def test_random_lists_match_oracle(sg, oracle, sizes, overlap):
    rng = np.random.default_rng(sum(sizes) + len(sizes))
    universe = pool(rng, max(sum(sizes), 1))
    shared = universe[: max(1, int(len(universe) * overlap / 2))]
    lists = []
    for n in sizes:
        n_sh = min(int(n * overlap), len(shared))
        own = universe[rng.choice(len(universe), n - n_sh, replace=False)] if n - n_sh else np.zeros((0, 2), np.uint64)
        sh = shared[rng.choice(len(shared), n_sh, replace=False)] if n_sh else np.zeros((0, 2), np.uint64)
        l = np.concatenate([own, sh])
        lists.append(l[rng.permutation(len(l))])
    for k, w in ((45.0, rng.random(len(sizes)).astype(np.float32) + 0.05), (60.0, np.ones(len(sizes), np.float32))):
        same(sg.RRFusion(k, w).fuse(lists), oracle.rrf_fuse(k, w, lists))


def test_weight_edge_cases(sg, oracle):
    rng = np.random.default_rng(3)
    lists = [pool(rng, 20), pool(rng, 30), pool(rng, 10)]
    lists[2][:5] = lists[0][:5]
    for w in ([0.0, 0.0, 0.0],            # sum <= 0 -> 1/len each (hybrid_search.rs:540-545)
              [1.0],                      # fewer weights than lists -> the rest use 1.0 (:562)
              [-1.0, 0.5, 0.25],          # negative sum
              [3.0, 1.0, 1.0, 7.0]):      # more weights than lists (normalised over all of them)
        same(sg.RRFusion(45.0, w).fuse(lists), oracle.rrf_fuse(45.0, w, lists))
    # the same id twice inside one list accumulates twice, in rank order (HashMap += semantics)
    dup = [np.concatenate([lists[0][:3], lists[0][:3]]), lists[1][:4]]
    same(sg.RRFusion(60.0, [0.5, 0.5]).fuse(dup), oracle.rrf_fuse(60.0, [0.5, 0.5], dup))


def test_limits(sg):
    rng = np.random.default_rng(4)
    with pytest.raises(sg.SgError):
        sg.RRFusion(60.0, [0.5, 0.5]).fuse([pool(rng, 2000), pool(rng, 49)])      # > 2048 entries
    ids, sc = sg.RRFusion(60.0, [0.5, 0.5]).fuse([])
    assert ids.shape == (0, 2) and sc.shape == (0,)
