"""Committed golden fixtures for the paths beside the NormalizedDotProduct scan (tests/golden/index_golden_v1.npz, made by
tests/golden/make_index_golden.py from the pinned oracle; inputs AND expected outputs are stored): the oracle must still
reproduce them on the CPU, the device must reproduce them bit for bit."""
import os

import numpy as np
import pytest


G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "index_golden_v1.npz"))
F = np.float32
NONE = 0xFFFFFFFF


def bits(a):
    return np.ascontiguousarray(a, F).view(np.uint32)


# ------------------------------------------------------------------ CPU: the oracle against the committed bytes
def test_oracle_reproduces_metric_golden(oracle):
    rows, qs = G["m_rows"], G["m_queries"]
    for mname, mid in (("euclidean", 1), ("cosine", 2)):
        for vname, var in (("scalar", 0), ("avx2", 1)):
            ids, dist, _, _ = oracle.brute_force_search_batch(rows, qs, 12, deleted=[9], metric=mid, variant=var)
            assert np.array_equal(ids, G[f"m_{mname}_{vname}_ids"]) and np.array_equal(bits(dist), G[f"m_{mname}_{vname}_dist_bits"])
    for vname, var in (("scalar", 0), ("avx2", 1)):
        pos, sc = oracle.top_k_similar(qs[0], rows, 12, variant=var)
        assert np.array_equal(pos, G[f"m_topk_{vname}_pos"]) and np.array_equal(bits(sc), G[f"m_topk_{vname}_score_bits"])


def test_oracle_reproduces_graph_golden(oracle):
    V = oracle.VamanaOracle(64, max_degree=12, search_list_size=30)
    V.set_graph(G["g_rows"], G["g_deg"], G["g_nbr"], int(G["g_medoid"][0]))
    for d in G["g_deleted"]:
        V.mark_deleted(int(d))
    for i, q in enumerate(G["g_queries"]):
        ids, dist, st = V.search(q, 10, stats=True)
        c = int(G["g_counts"][i])
        assert c == len(ids) and np.array_equal(ids, G["g_ids"][i, :c]) and np.array_equal(bits(dist), G["g_dist_bits"][i, :c])
        assert np.array_equal(st, G["g_stats"][i])


def test_oracle_reproduces_pq_and_spann_golden(oracle):
    pq = oracle.PQOracle(32, num_centroids=32, kmeans_iterations=3).train(G["p_train"], orders=G["p_orders"])
    assert np.array_equal(bits(pq.centroids), G["p_centroid_bits"])
    pq.add(G["p_ids"], G["p_train"])
    assert np.array_equal(pq.codes, G["p_codes"]) and np.array_equal(bits(pq.table(G["p_queries"][0])), G["p_table_bits"])
    for i, q in enumerate(G["p_queries"]):
        ids, dist = pq.search(q, 9)
        assert np.array_equal(ids, G["p_res_ids"][i]) and np.array_equal(bits(dist), G["p_res_dist_bits"][i])
    sp = oracle.SpannOracle(64, num_probes=5, kmeans_iterations=4, pq_iterations=2).build(
        G["s_rows"], order=G["s_order"], pq_orders=G["s_pq_orders"])
    assert np.array_equal(bits(sp.centroids), G["s_centroid_bits"]) and np.array_equal(bits(sp.pq.centroids), G["s_pq_centroid_bits"])
    off, ids, codes = sp.csr()
    assert np.array_equal(off, G["s_off"]) and np.array_equal(ids, G["s_ids"]) and np.array_equal(codes, G["s_codes"])
    for i, q in enumerate(G["s_queries"]):
        ri, rd, pr = sp.search(q, 8, with_probes=True)
        assert np.array_equal(ri, G["s_res_ids"][i]) and np.array_equal(bits(rd), G["s_res_dist_bits"][i])
        assert np.array_equal(pr, G["s_probes"][i])


# ------------------------------------------------------------------ GPU: the device against the committed bytes
@pytest.mark.gpu
@pytest.mark.parametrize("order", ["scalar", "avx2"])
def test_gpu_reproduces_metric_golden(sg, order):
    rows, qs = G["m_rows"], G["m_queries"]
    for metric in ("euclidean", "cosine"):
        st = sg.GpuVectorStore(64, "f32", metric=metric)
        st.set_order(order)
        st.add_vectors(rows)
        st.mark_deleted(9)
        ids, dist, _ = st.search_batch(qs, 12)
        assert np.array_equal(ids, G[f"m_{metric}_{order}_ids"]) and np.array_equal(bits(dist), G[f"m_{metric}_{order}_dist_bits"])
        st.close()
    pos, sc = sg.top_k_similar(qs[0], rows, 12, order=order)
    assert np.array_equal(pos, G[f"m_topk_{order}_pos"])
    assert np.array_equal(sc, G[f"m_topk_{order}_score_bits"].view(F))          # == on values: -0.0 and +0.0 compare equal


@pytest.mark.gpu
def test_gpu_reproduces_graph_golden(sg):
    st = sg.GpuVectorStore(64, "f32")
    st.add_vectors(G["g_rows"])
    st.set_graph(G["g_deg"], G["g_nbr"], int(G["g_medoid"][0]), 12)
    for d in G["g_deleted"]:
        st.mark_deleted(int(d))
    ids, dist, cnt, stats = st.search_ann_batch(G["g_queries"], 10)
    assert np.array_equal(cnt, G["g_counts"]) and np.array_equal(ids, G["g_ids"])
    assert np.array_equal(bits(dist), G["g_dist_bits"]) and np.array_equal(stats, G["g_stats"])
    st.close()


@pytest.mark.gpu
def test_gpu_reproduces_pq_and_spann_golden(sg):
    pq = sg.GpuProductQuantizer.train(G["p_train"], G["p_orders"], num_centroids=32, kmeans_iterations=3)
    assert np.array_equal(bits(pq.centroids), G["p_centroid_bits"])
    assert np.array_equal(pq.encode(G["p_train"]), G["p_codes"])
    assert np.array_equal(bits(pq.build_distance_table(G["p_queries"][0])), G["p_table_bits"])
    pq.add(G["p_ids"], G["p_train"])
    ids, dist, cnt = pq.search_batch(G["p_queries"], 9)
    assert (cnt == 9).all() and np.array_equal(ids, G["p_res_ids"]) and np.array_equal(bits(dist), G["p_res_dist_bits"])
    pq.close()
    sp = sg.GpuSpannIndex.build(G["s_rows"], G["s_order"], G["s_pq_orders"], num_probes=5, kmeans_iterations=4, pq_iterations=2)
    cent, pqc, off, sids, codes = sp.arrays(G["s_pq_centroid_bits"].shape)
    assert np.array_equal(bits(cent), G["s_centroid_bits"]) and np.array_equal(bits(pqc), G["s_pq_centroid_bits"])
    assert np.array_equal(off, G["s_off"]) and np.array_equal(sids, G["s_ids"]) and np.array_equal(codes, G["s_codes"])
    ids, dist, cnt = sp.search_batch(G["s_queries"], 8)
    assert (cnt == 8).all() and np.array_equal(ids, G["s_res_ids"]) and np.array_equal(bits(dist), G["s_res_dist_bits"])
    sp.close()
