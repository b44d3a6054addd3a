"""Generates tests/golden/scan_golden_v1.npz from the CPU oracle (run from the repo root).

The reference holds NO golden top-k lists or golden embeddings for this path, so these
fixtures are produced by the oracle -- which is itself pinned to every known-answer unit test the reference
has (tests/test_oracle_pins.py).  They freeze the oracle's answers so that (a) an accidental change to the
oracle's arithmetic is caught on CPU, and (b) the GPU path is checked against committed bytes as well as
against the live oracle.  Inputs are regenerated from the seed; only outputs are stored."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402

SEED, N, DIM, B, K = 23821, 3000, 384, 12, 25
DELETED = [0, 17, 2999]


def build():
    out = {}
    for tag, f16 in (("f16", True), ("f32", False)):
        seen = o.synth_corpus(SEED, 0, N, DIM, f16_round=f16)
        q, near = o.synth_queries(SEED, N, B, DIM)
        for variant, vname in ((0, "scalar"), (1, "avx2")):
            ids, dist, cnt, _ = o.brute_force_search_batch(seen, q, K, deleted=DELETED, variant=variant)
            out[f"ids_{tag}_{vname}"] = ids
            out[f"dist_bits_{tag}_{vname}"] = dist.view(np.uint32)
        out[f"row_checksum_{tag}"] = np.array([int(seen.view(np.uint32).astype(np.uint64).sum() & 0xFFFFFFFFFFFFFFFF)], np.uint64)
    out["query_checksum"] = np.array([int(q.view(np.uint32).astype(np.uint64).sum())], np.uint64)
    now = 1_790_000_000
    rng = np.random.default_rng(SEED)
    m = 100
    feats = dict(semantic=rng.random(m).astype(np.float32), entity=(rng.random(m) * (rng.random(m) < 0.3)).astype(np.float32),
                 tag=(rng.random(m) * (rng.random(m) < 0.3)).astype(np.float32), importance=rng.random(m).astype(np.float32),
                 momentum=rng.uniform(-1, 1, m).astype(np.float32), access_count=np.minimum(rng.geometric(0.3, m), 64).astype(np.uint32),
                 graph=rng.random(m).astype(np.float32), created_at=(now - rng.integers(0, 72 * 3600, m)).astype(np.int64))
    order, score, allsc = o.fuse_rank(feats, now, max_results=20)
    out.update({f"fuse_{k}": v for k, v in feats.items()})
    out["fuse_order"], out["fuse_score_bits"], out["fuse_all_bits"] = order, score.view(np.uint32), allsc.view(np.uint32)
    out["meta"] = np.array([SEED, N, DIM, B, K, now] + DELETED, np.int64)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "scan_golden_v1.npz"), **build())
    print("wrote scan_golden_v1.npz")
