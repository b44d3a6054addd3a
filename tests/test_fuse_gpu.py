"""GPU parity for relevance fusion.  Floating point: fused scores
must agree with the oracle within 1e-6 absolute (exp/log2 are evaluated in f64 and rounded once
on the GPU vs glibc expf/log2f on the CPU); the ranking must be identical wherever the oracle's
own scores are separated by more than that tolerance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F = np.float32
TOL = 1e-6


def make_feats(n, seed, now):
    rng = np.random.default_rng(seed)
    ent = np.where(rng.random(n) < 0.3, rng.random(n), 0.0)
    tag = np.where(rng.random(n) < 0.3, rng.random(n), 0.0)
    return dict(semantic=rng.random(n).astype(F), entity=ent.astype(F), tag=tag.astype(F),
                importance=rng.random(n).astype(F), momentum=rng.uniform(-1, 1, n).astype(F),
                access_count=np.minimum(rng.geometric(0.3, n), 64).astype(np.uint32),
                graph=rng.random(n).astype(F),
                created_at=(now - rng.integers(0, 72 * 3600, n)).astype(np.int64))


@pytest.mark.parametrize("n", [1, 7, 100, 120, 1000])
def test_fuse_matches_oracle(sg, oracle, n):
    now = 1_790_000_000
    feats = make_feats(n, n, now)
    order, score, allsc = sg.fuse_topk(feats, now, max_results=n)
    o_order, o_score, o_all = oracle.fuse_rank(feats, now, max_results=n)
    assert np.abs(allsc - o_all).max() <= TOL
    assert len(order) == len(o_order) or abs(len(order) - len(o_order)) <= int((np.abs(o_all - 0.25) <= TOL).sum())
    exact = (allsc.view(np.uint32) == o_all.view(np.uint32)).mean()
    assert exact > 0.95                                 # f64-rounded exp == glibc expf almost always
    if np.array_equal(allsc.view(np.uint32), o_all.view(np.uint32)):
        assert np.array_equal(order, o_order) and np.array_equal(score.view(np.uint32), o_score.view(np.uint32))
    else:                                               # ranking equal up to tolerance-level near-ties
        m = min(len(order), len(o_order))
        assert np.abs(score[:m] - o_score[:m]).max() <= TOL


def test_fuse_reference_pins(sg):
    """relevance.rs:1960-1979 and :2044-2066 through the GPU kernel."""
    now = 1_790_000_000
    old = now - 100 * 3600

    def one(sem, ent, tag, imp, mom, acc, gr):
        f = dict(semantic=[sem], entity=[ent], tag=[tag], importance=[imp], momentum=[mom],
                 access_count=[acc], graph=[gr], created_at=[old])
        _, _, allsc = sg.fuse_topk(f, now, max_results=1, graph_boost=1.0)
        return float(allsc[0])
    assert one(0.9, 0.9, 0.9, 0.9, 0.9, 16, 0.9) > 0.8
    assert one(0.1, 0.1, 0.1, 0.1, -0.9, 0, 0.1) < 0.3
    assert 0.2 < one(0.9, 0.1, 0.5, 0.7, 0.0, 2, 0.5) < 0.8
    for bad in (float("nan"), float("inf"), float("-inf")):
        assert np.isfinite(one(bad, 0.5, 0.5, 0.5, 0.0, 1, 0.5))
        assert np.isfinite(one(0.5, bad, 0.5, 0.5, 0.0, 1, 0.5))


def test_fuse_sort_threshold_truncate(sg, oracle):
    """relevance.rs:905-918 tie-breaks: created_at desc, then position asc; >= 0.25; truncate."""
    now = 2_000_000_000
    n = 6
    feats = dict(semantic=[0.9, 0.9, 0.1, 0.9, 0.5, 0.9], entity=[0.9, 0.9, 0.0, 0.9, 0.0, 0.9],
                 tag=[0.5] * n, importance=[0.5] * n, momentum=[0.5, 0.5, -0.9, 0.5, 0.0, 0.5],
                 access_count=[4, 4, 0, 4, 1, 4], graph=[0.5, 0.5, 0.1, 0.5, 0.5, 0.5],
                 created_at=[now - 100 * 3600, now - 90 * 3600, now - 100 * 3600, now - 100 * 3600,
                             now - 100 * 3600, now - 90 * 3600])
    order, score, _ = sg.fuse_topk(feats, now, max_results=10)
    o_order, _, _ = oracle.fuse_rank(feats, now, max_results=10)
    assert list(order[:4]) == [1, 5, 0, 3] and list(order) == list(o_order)
    assert 2 not in order and (score >= 0.25).all()
    order2, _, _ = sg.fuse_topk(feats, now, max_results=2)
    assert list(order2) == [1, 5]
