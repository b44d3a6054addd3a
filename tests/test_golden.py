"""Committed golden fixtures (tests/golden/scan_golden_v1.npz, made by tests/golden/make_golden.py from the
pinned oracle): the oracle must still reproduce them on CPU; the GPU path must reproduce them bit for bit."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scan_golden_v1.npz"))
SEED, N, DIM, B, K, NOW = (int(x) for x in G["meta"][:6])
DELETED = [int(x) for x in G["meta"][6:]]


def test_oracle_reproduces_golden(oracle):
    for tag, f16 in (("f16", True), ("f32", False)):
        seen = oracle.synth_corpus(SEED, 0, N, DIM, f16_round=f16)
        assert int(seen.view(np.uint32).astype(np.uint64).sum()) == int(G[f"row_checksum_{tag}"][0])
        q, _ = oracle.synth_queries(SEED, N, B, DIM)
        for variant, vname in ((0, "scalar"), (1, "avx2")):
            ids, dist, _, _ = oracle.brute_force_search_batch(seen, q, K, deleted=DELETED, variant=variant)
            assert np.array_equal(ids, G[f"ids_{tag}_{vname}"])
            assert np.array_equal(dist.view(np.uint32), G[f"dist_bits_{tag}_{vname}"])
    feats = {k[5:]: G[k] for k in G.files if k.startswith("fuse_") and k not in ("fuse_order", "fuse_score_bits", "fuse_all_bits")}
    order, score, allsc = oracle.fuse_rank(feats, NOW, max_results=20)
    assert np.array_equal(order, G["fuse_order"]) and np.array_equal(score.view(np.uint32), G["fuse_score_bits"])
    assert np.array_equal(allsc.view(np.uint32), G["fuse_all_bits"])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["f16", "f32"])
@pytest.mark.parametrize("order", ["scalar", "avx2"])
def test_gpu_reproduces_golden(sg, oracle, dtype, order):
    st = sg.GpuVectorStore(DIM, dtype)
    st.append_synthetic(SEED, 0, N)
    st.set_order(order)
    for d in DELETED:
        st.mark_deleted(d)
    q, _ = oracle.synth_queries(SEED, N, B, DIM)
    for b in (1, 4, B):                                   # exact sweep (b <= 4) and the 12-query batch
        ids, dist, cnt = st.search_batch(q[:b], K)
        assert np.array_equal(ids, G[f"ids_{dtype}_{order}"][:b])
        assert np.array_equal(dist.view(np.uint32), G[f"dist_bits_{dtype}_{order}"][:b])
    st.close()


@pytest.mark.gpu
def test_gpu_fusion_matches_golden_within_tolerance(sg):
    feats = {k[5:]: G[k] for k in G.files if k.startswith("fuse_") and k not in ("fuse_order", "fuse_score_bits", "fuse_all_bits")}
    order, score, allsc = sg.fuse_topk(feats, NOW, max_results=20)
    assert np.abs(allsc - G["fuse_all_bits"].view(np.float32)).max() <= 1e-6
    assert list(order) == list(G["fuse_order"]) or np.abs(score - G["fuse_score_bits"].view(np.float32)).max() <= 1e-6
