"""Makes tests/golden/index_golden_v1.npz from the pinned CPU oracle: small committed inputs + expected outputs for the
paths beside the NormalizedDotProduct scan (metrics / top_k_similar, Vamana graph search, Product Quantization, SPANN).
Run from the repo root:  python tests/golden/make_index_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402

F = np.float32
out = {}
rng = np.random.default_rng(20261017)

# ---- metrics + top_k_similar: 600 x 64 unnormalised rows, 6 queries ----
rows = (rng.standard_normal((600, 64)) * rng.uniform(0.2, 4.0, (600, 1))).astype(F)
rows[77] = 0.0
qs = (rng.standard_normal((6, 64)) * 1.5).astype(F)
qs[0] = rows[5] * F(1.25)
out["m_rows"], out["m_queries"] = rows, qs
for mname, mid in (("euclidean", 1), ("cosine", 2)):
    for vname, var in (("scalar", 0), ("avx2", 1)):
        ids, dist, _, _ = o.brute_force_search_batch(rows, qs, 12, deleted=[9], metric=mid, variant=var)
        out[f"m_{mname}_{vname}_ids"], out[f"m_{mname}_{vname}_dist_bits"] = ids, dist.view(np.uint32)
for vname, var in (("scalar", 0), ("avx2", 1)):
    pos, sc = o.top_k_similar(qs[0], rows, 12, variant=var)
    out[f"m_topk_{vname}_pos"], out[f"m_topk_{vname}_score_bits"] = pos, sc.view(np.uint32)

# ---- Vamana: 700 x 64 rows on a manifold, oracle-built graph (R 12, L 30), 8 queries ----
gx = o.manifold_corpus(11, 700, 64, latent=6)
V = o.VamanaOracle(64, max_degree=12, search_list_size=30)
V.build(gx, seed=5)
gq = (gx[rng.integers(0, 700, 8)] + F(0.03) * rng.standard_normal((8, 64))).astype(F)
out["g_rows"], out["g_deg"], out["g_nbr"], out["g_medoid"], out["g_queries"] = gx, V.deg, V.nbr[:, :12], np.array([V.medoid]), gq
V.mark_deleted(3)
V.mark_deleted(int(V.search(gq[0], 1)[0][0]))
out["g_deleted"] = np.array(sorted(V.deleted), np.uint32)
res = [V.search(gq[i], 10, stats=True) for i in range(8)]
out["g_ids"] = np.stack([np.pad(r[0], (0, 10 - len(r[0])), constant_values=0xFFFFFFFF) for r in res])
out["g_dist_bits"] = np.stack([np.pad(r[1], (0, 10 - len(r[1])), constant_values=np.inf) for r in res]).view(np.uint32)
out["g_stats"] = np.stack([r[2] for r in res])
out["g_counts"] = np.array([len(r[0]) for r in res], np.uint32)

# ---- PQ: dim 32 (4 subvectors), 32 centroids, 400 training rows, 3 iterations ----
pv = rng.standard_normal((400, 32)).astype(F)
pq = o.PQOracle(32, num_centroids=32, kmeans_iterations=3)
orders = pq.shuffled_orders(400, 9)
pq.train(pv, orders=orders)
pids = rng.permutation(np.arange(1000, 1400)).astype(np.uint32)
pq.add(pids, pv)
pqq = rng.standard_normal((4, 32)).astype(F)
out["p_train"], out["p_orders"], out["p_centroid_bits"], out["p_ids"], out["p_codes"], out["p_queries"] = \
    pv, orders, pq.centroids.view(np.uint32), pids, pq.codes, pqq
out["p_table_bits"] = pq.table(pqq[0]).view(np.uint32)
pr = [pq.search(pqq[i], 9) for i in range(4)]
out["p_res_ids"], out["p_res_dist_bits"] = np.stack([r[0] for r in pr]), np.stack([r[1] for r in pr]).view(np.uint32)

# ---- SPANN: 900 x 64 unit rows, 30 partitions, 5 probes ----
sv = rng.random((900, 64)).astype(F)
sv = np.ascontiguousarray(sv / np.linalg.norm(sv, axis=1, keepdims=True), F)
sp = o.SpannOracle(64, num_probes=5, kmeans_iterations=4, pq_iterations=2).build(sv, seed=2)
off, sids, scodes = sp.csr()
sq = np.ascontiguousarray(sv[rng.integers(0, 900, 5)] + F(0.01) * rng.standard_normal((5, 64)).astype(F), F)
out["s_rows"], out["s_order"], out["s_pq_orders"] = sv, sp.order, sp.pq_orders
out["s_centroid_bits"], out["s_pq_centroid_bits"] = sp.centroids.view(np.uint32), sp.pq.centroids.view(np.uint32)
out["s_off"], out["s_ids"], out["s_codes"], out["s_queries"] = off, sids, scodes, sq
sr = [sp.search(sq[i], 8, with_probes=True) for i in range(5)]
out["s_res_ids"], out["s_res_dist_bits"] = np.stack([r[0] for r in sr]), np.stack([r[1] for r in sr]).view(np.uint32)
out["s_probes"] = np.stack([r[2] for r in sr])

path = os.path.join(ROOT, "tests", "golden", "index_golden_v1.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes")
