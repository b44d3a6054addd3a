"""ctypes binding of the CPU oracle (oracle/shodh_oracle.c).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by shodh_memory_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_SRC = os.path.join(_HERE, "shodh_oracle.c")
_SRCS = [_SRC, os.path.join(_HERE, "vamana_oracle.c"), os.path.join(_HERE, "pq_oracle.c"), os.path.join(_HERE, "spann_oracle.c"),
         os.path.join(_HERE, "build.sh")]

DEFAULT_WEIGHTS = np.array([0.18, 0.17, 0.05, 0.05, 0.28, 0.14, 0.13], dtype=np.float32)

METRIC_NEG_DOT, METRIC_EUCLIDEAN, METRIC_COSINE = 0, 1, 2
VARIANT_SCALAR, VARIANT_AVX2 = 0, 1


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(p) for p in _SRCS):
        subprocess.check_call(["sh", os.path.join(_HERE, "build.sh")])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        f32p, u32p, u64p, i64p = (C.POINTER(C.c_float), C.POINTER(C.c_uint32),
                                  C.POINTER(C.c_uint64), C.POINTER(C.c_int64))
        L.so_dot.restype = C.c_float
        L.so_dot.argtypes = [f32p, f32p, C.c_uint32, C.c_int]
        L.so_euclidean_squared.restype = C.c_float
        L.so_euclidean_squared.argtypes = [f32p, f32p, C.c_uint32]
        L.so_euclidean_squared_avx2.restype = C.c_float
        L.so_euclidean_squared_avx2.argtypes = [f32p, f32p, C.c_uint32]
        L.so_l2_norm_squared_avx2.restype = C.c_float
        L.so_l2_norm_squared_avx2.argtypes = [f32p, C.c_uint32]
        L.so_cosine_similarity_inline_v.restype = C.c_float
        L.so_cosine_similarity_inline_v.argtypes = [f32p, f32p, C.c_uint32, C.c_int]
        L.so_distance.restype = C.c_float
        L.so_distance.argtypes = [f32p, f32p, C.c_uint32, C.c_int, C.c_int]
        L.so_l2_norm.restype = C.c_float
        L.so_l2_norm.argtypes = [f32p, C.c_uint32]
        L.so_l2_norm_squared.restype = C.c_float
        L.so_l2_norm_squared.argtypes = [f32p, C.c_uint32]
        L.so_normalized_distance.restype = C.c_float
        L.so_normalized_distance.argtypes = [f32p, f32p, C.c_uint32, C.c_int]
        L.so_cosine_similarity.restype = C.c_float
        L.so_cosine_similarity.argtypes = [f32p, C.c_uint32, f32p, C.c_uint32, C.c_int]
        L.so_cosine_similarity_inline.restype = C.c_float
        L.so_cosine_similarity_inline.argtypes = [f32p, f32p, C.c_uint32]
        L.so_normalize_inplace.restype = None
        L.so_normalize_inplace.argtypes = [f32p, C.c_uint32]
        L.so_total_order_key.restype = C.c_uint32
        L.so_total_order_key.argtypes = [C.c_float]
        L.so_brute_force_search.restype = C.c_uint32
        L.so_brute_force_search.argtypes = [f32p, C.c_uint64, C.c_uint32, u32p, f32p, C.c_uint32,
                                            C.c_int, C.c_int, C.c_int, u32p, f32p]
        L.so_brute_force_search_batch.restype = C.c_int
        L.so_brute_force_search_batch.argtypes = [f32p, C.c_uint64, C.c_uint32, u32p, f32p,
                                                  C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                                  C.c_int, u32p, f32p, u32p]
        L.so_score_rows.restype = None
        L.so_score_rows.argtypes = [f32p, C.c_uint32, f32p, u32p, C.c_uint32, C.c_int, C.c_int, f32p]
        L.so_top_k_similar.restype = C.c_uint32
        L.so_top_k_similar.argtypes = [f32p, C.c_uint32, f32p, C.c_uint64, C.c_uint32, C.c_uint32,
                                       C.c_int, u32p, f32p]
        L.so_search_ids_post.restype = C.c_uint32
        L.so_search_ids_post.argtypes = [u32p, f32p, C.c_uint32, u64p, C.c_uint64, C.c_uint32,
                                         u64p, f32p]
        L.so_calibrate_score.restype = C.c_float
        L.so_calibrate_score.argtypes = [C.c_float]
        L.so_fuse_scores_full.restype = C.c_float
        L.so_fuse_scores_full.argtypes = [f32p] + [C.c_float] * 5 + [C.c_uint32, C.c_float]
        L.so_apply_recency_boost.restype = C.c_float
        L.so_apply_recency_boost.argtypes = [C.c_float, C.c_int64, C.c_int64, C.c_uint64, C.c_float]
        L.so_fuse_rank.restype = C.c_uint32
        L.so_fuse_rank.argtypes = [f32p] * 6 + [u32p, f32p, i64p, C.c_uint32, C.c_int64, C.c_uint64,
                                                C.c_float, C.c_float, C.c_float, C.c_uint32, u32p,
                                                f32p, f32p]
        L.so_pool_finalize.restype = None
        L.so_pool_finalize.argtypes = [f32p, i64p, C.c_uint32, C.c_uint32, f32p]
        L.so_synth_corpus.restype = None
        L.so_synth_corpus.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, f32p]
        L.so_synth_queries.restype = None
        L.so_synth_queries.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, f32p, u64p]
        L.so_round_f16.restype = None
        L.so_round_f16.argtypes = [f32p, f32p, C.c_uint64]
        L.so_max_threads.restype = C.c_int
        L.so_max_threads.argtypes = []
        # vamana_oracle.c
        L.so_greedy_search.restype = C.c_uint32
        L.so_greedy_search.argtypes = [f32p, C.c_uint64, C.c_uint32, u32p, u32p, C.c_uint32, C.c_uint64, f32p,
                                       C.c_uint32, C.c_uint32, C.c_int, C.c_int, u32p, f32p, u64p]
        L.so_vamana_search.restype = C.c_uint32
        L.so_vamana_search.argtypes = [f32p, C.c_uint64, C.c_uint32, u32p, u32p, C.c_uint32, C.c_uint32, u32p,
                                       C.c_uint64, f32p, C.c_uint32, C.c_int, C.c_int, u32p, f32p, u64p]
        L.so_robust_prune.restype = C.c_uint32
        L.so_robust_prune.argtypes = [f32p, C.c_uint32, C.c_uint32, u32p, f32p, C.c_uint32, C.c_uint32, C.c_float,
                                      C.c_int, C.c_int, u32p]
        L.so_find_medoid.restype = C.c_uint32
        L.so_find_medoid.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_int, C.c_int]
        L.so_vamana_init_graph.restype = None
        L.so_vamana_init_graph.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, u32p, u32p]
        L.so_vamana_build.restype = C.c_uint32
        L.so_vamana_build.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_int,
                                      C.c_int, C.c_uint32, C.c_uint32, u32p, u32p]
        L.so_vamana_add_vector.restype = None
        L.so_vamana_add_vector.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint32,
                                           C.c_uint32, u32p, u32p]
        # pq_oracle.c
        u8p = C.POINTER(C.c_uint8)
        L.so_pq_train.restype = C.c_uint32
        L.so_pq_train.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u32p, f32p]
        L.so_pq_encode.restype = None
        L.so_pq_encode.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_uint32, f32p, C.c_uint64, u8p]
        L.so_pq_decode.restype = None
        L.so_pq_decode.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_uint32, u8p, C.c_uint64, f32p]
        L.so_pq_table.restype = None
        L.so_pq_table.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_uint32, f32p, f32p]
        L.so_pq_distance_with_table.restype = C.c_float
        L.so_pq_distance_with_table.argtypes = [f32p, C.c_uint32, C.c_uint32, u8p]
        L.so_pq_asymmetric_distance.restype = C.c_float
        L.so_pq_asymmetric_distance.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_uint32, f32p, u8p]
        L.so_pq_search.restype = C.c_uint32
        L.so_pq_search.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_uint32, u32p, u8p, C.c_uint64, f32p, C.c_uint32,
                                   u32p, f32p]
        # spann_oracle.c
        L.so_spann_distance.restype = C.c_float
        L.so_spann_distance.argtypes = [f32p, f32p, C.c_uint32, C.c_int]
        L.so_spann_nearest_centroid.restype = C.c_uint32
        L.so_spann_nearest_centroid.argtypes = [f32p, f32p, C.c_uint32, C.c_uint32, C.c_int]
        L.so_spann_kmeans.restype = C.c_uint32
        L.so_spann_kmeans.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, u32p, f32p]
        L.so_spann_assign.restype = None
        L.so_spann_assign.argtypes = [f32p, C.c_uint64, C.c_uint32, f32p, C.c_uint32, C.c_int, u32p]
        L.so_spann_search.restype = C.c_uint32
        L.so_spann_search.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, u64p, u32p, u8p, f32p, C.c_uint32,
                                      C.c_uint32, C.c_uint32, f32p, C.c_uint32, u32p, f32p, u32p]
        L.so_recall_at_k.restype = C.c_double
        L.so_recall_at_k.argtypes = [u32p, u32p, u32p, u32p, C.c_uint32, C.c_uint32]
        _lib = L
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def dot(a, b, variant=VARIANT_SCALAR) -> float:
    a, b = _f32(a), _f32(b)
    assert a.shape == b.shape
    return float(lib().so_dot(_p(a, C.c_float), _p(b, C.c_float), a.size, variant))


def euclidean_squared(a, b, variant=VARIANT_SCALAR) -> float:
    a, b = _f32(a), _f32(b)
    fn = lib().so_euclidean_squared_avx2 if variant == VARIANT_AVX2 else lib().so_euclidean_squared
    return float(fn(_p(a, C.c_float), _p(b, C.c_float), a.size))


def l2_norm_squared(a, variant=VARIANT_SCALAR) -> float:
    a = _f32(a)
    fn = lib().so_l2_norm_squared_avx2 if variant == VARIANT_AVX2 else lib().so_l2_norm_squared
    return float(fn(_p(a, C.c_float), a.size))


def distance(a, b, metric=METRIC_NEG_DOT, variant=VARIANT_SCALAR) -> float:
    """VamanaIndex::distance (vamana.rs:755-761)."""
    a, b = _f32(a), _f32(b)
    return float(lib().so_distance(_p(a, C.c_float), _p(b, C.c_float), a.size, metric, variant))


def l2_norm(a) -> float:
    a = _f32(a)
    return float(lib().so_l2_norm(_p(a, C.c_float), a.size))


def normalized_distance(a, b, variant=VARIANT_SCALAR) -> float:
    a, b = _f32(a), _f32(b)
    return float(lib().so_normalized_distance(_p(a, C.c_float), _p(b, C.c_float), a.size, variant))


def cosine_similarity(a, b, variant=VARIANT_SCALAR) -> float:
    a, b = _f32(a), _f32(b)
    return float(lib().so_cosine_similarity(_p(a, C.c_float), a.size, _p(b, C.c_float), b.size, variant))


def cosine_similarity_inline(a, b, variant=VARIANT_SCALAR) -> float:
    a, b = _f32(a), _f32(b)
    return float(lib().so_cosine_similarity_inline_v(_p(a, C.c_float), _p(b, C.c_float), a.size, variant))


def normalize(a) -> np.ndarray:
    a = _f32(a).copy()
    lib().so_normalize_inplace(_p(a, C.c_float), a.size)
    return a


def total_order_key(x: float) -> int:
    return int(lib().so_total_order_key(C.c_float(x)))


def deleted_bitmap(n: int, deleted_ids) -> np.ndarray:
    bm = np.zeros((n + 31) // 32 + 1, dtype=np.uint32)
    for i in deleted_ids:
        bm[i >> 5] |= np.uint32(1 << (i & 31))
    return bm


def brute_force_search(corpus, query, k, deleted=None, metric=METRIC_NEG_DOT,
                       variant=VARIANT_SCALAR, as_written=False):
    """VamanaIndex::brute_force_search (vamana.rs:1167-1188) -> (ids u32[r], dist f32[r])."""
    corpus = _f32(corpus)
    n, dim = (corpus.shape if corpus.ndim == 2 else (0, 0))
    query = _f32(query)
    ids = np.zeros(max(k, 1), dtype=np.uint32)
    dist = np.zeros(max(k, 1), dtype=np.float32)
    bm = None if deleted is None else deleted_bitmap(n, deleted)
    r = lib().so_brute_force_search(_p(corpus, C.c_float), n, dim,
                                    None if bm is None else _p(bm, C.c_uint32),
                                    _p(query, C.c_float), k, metric, variant, int(as_written),
                                    _p(ids, C.c_uint32), _p(dist, C.c_float))
    return ids[:r].copy(), dist[:r].copy()


def brute_force_search_batch(corpus, queries, k, deleted=None, metric=METRIC_NEG_DOT,
                             variant=VARIANT_SCALAR, as_written=False, threads=0):
    """Batch of queries, one per host thread -> (ids [b,k], dist [b,k], counts [b], threads_used)."""
    corpus, queries = _f32(corpus), _f32(queries)
    n, dim = corpus.shape
    b = queries.shape[0]
    ids = np.zeros((b, max(k, 1)), dtype=np.uint32)
    dist = np.zeros((b, max(k, 1)), dtype=np.float32)
    counts = np.zeros(b, dtype=np.uint32)
    bm = None if deleted is None else deleted_bitmap(n, deleted)
    used = lib().so_brute_force_search_batch(_p(corpus, C.c_float), n, dim,
                                             None if bm is None else _p(bm, C.c_uint32),
                                             _p(queries, C.c_float), b, k, metric, variant,
                                             int(as_written), threads, _p(ids, C.c_uint32),
                                             _p(dist, C.c_float), _p(counts, C.c_uint32))
    return ids, dist, counts, used


def order_keys(dist: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """u64 keys whose unsigned order is (dist f32::total_cmp asc, id asc) (vamana.rs:1184)."""
    b = np.ascontiguousarray(dist, np.float32).view(np.uint32).astype(np.uint64)
    k32 = np.where(b >> np.uint64(31), (~b) & np.uint64(0xFFFFFFFF), b | np.uint64(0x80000000))
    return (k32 << np.uint64(32)) | np.ascontiguousarray(ids, np.uint32).astype(np.uint64)


def brute_force_search_synth(seed, n_total, queries, k, deleted=None, dim=384, chunk=1_000_000,
                             variant=VARIANT_SCALAR, threads=0):
    """brute_force_search over the synthetic f16-rounded corpus rows [0, n_total) WITHOUT holding it in memory:
    the corpus is generated chunk by chunk, each chunk's top-k (global ids) is taken with so_brute_force_search_batch
    and the per-chunk lists are merged with the same (dist total_cmp, id) rule -- identical to sorting all rows at
    once, because the order is total.  -> (ids [b,k] u32, dist [b,k] f32, counts [b])."""
    queries = _f32(queries)
    b = queries.shape[0]
    dele = sorted(deleted or [])
    keys = [[] for _ in range(b)]
    for lo in range(0, n_total, chunk):
        m = min(chunk, n_total - lo)
        c = synth_corpus(seed, lo, m, dim, f16_round=True)
        dl = [d - lo for d in dele if lo <= d < lo + m]
        ids, dist, cnt, _ = brute_force_search_batch(c, queries, k, deleted=dl or None, variant=variant, threads=threads)
        for q in range(b):
            r = int(cnt[q])
            keys[q].append(order_keys(dist[q, :r], ids[q, :r].astype(np.uint64) + lo))
        del c
    out_ids = np.full((b, k), 0xFFFFFFFF, np.uint32)
    out_dist = np.full((b, k), np.inf, np.float32)
    counts = np.zeros(b, np.uint32)
    for q in range(b):
        kk = np.sort(np.concatenate(keys[q]))[:k]
        counts[q] = kk.size
        out_ids[q, :kk.size] = (kk & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        k32 = (kk >> np.uint64(32)).astype(np.uint32)
        bits = np.where(k32 & np.uint32(0x80000000), k32 ^ np.uint32(0x80000000), ~k32)
        out_dist[q, :kk.size] = bits.astype(np.uint32).view(np.float32)
    return out_ids, out_dist, counts


def score_rows(corpus, query, ids, metric=METRIC_NEG_DOT, variant=VARIANT_SCALAR) -> np.ndarray:
    corpus, query = _f32(corpus), _f32(query)
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    out = np.zeros(ids.size, dtype=np.float32)
    lib().so_score_rows(_p(corpus, C.c_float), corpus.shape[1], _p(query, C.c_float),
                        _p(ids, C.c_uint32), ids.size, metric, variant, _p(out, C.c_float))
    return out


def top_k_similar(query, candidates, k, variant=VARIANT_SCALAR):
    """similarity.rs::top_k_similar -> (positions, scores)."""
    query = _f32(query)
    candidates = _f32(candidates)
    if candidates.size == 0:
        return np.zeros(0, np.uint32), np.zeros(0, np.float32)
    n, dim = candidates.shape
    pos = np.zeros(max(k, 1), np.uint32)
    sc = np.zeros(max(k, 1), np.float32)
    r = lib().so_top_k_similar(_p(query, C.c_float), query.size, _p(candidates, C.c_float), n, dim,
                               k, variant, _p(pos, C.c_uint32), _p(sc, C.c_float))
    return pos[:r].copy(), sc[:r].copy()


def search_ids_post(vec_ids, dists, vec2mem, limit):
    """retrieval.rs:927-961 -> (mem ids [r,2] u64, similarities [r])."""
    vec_ids = np.ascontiguousarray(vec_ids, np.uint32)
    dists = _f32(dists)
    vec2mem = np.ascontiguousarray(vec2mem, np.uint64)
    mem = np.zeros((max(limit, 1), 2), np.uint64)
    sim = np.zeros(max(limit, 1), np.float32)
    r = lib().so_search_ids_post(_p(vec_ids, C.c_uint32), _p(dists, C.c_float), vec_ids.size,
                                 _p(vec2mem, C.c_uint64), vec2mem.shape[0], limit,
                                 _p(mem, C.c_uint64), _p(sim, C.c_float))
    return mem[:r].copy(), sim[:r].copy()


def calibrate_score(x: float) -> float:
    return float(lib().so_calibrate_score(C.c_float(x)))


def fuse_scores_full(sem, ent, tag, imp, mom, access, graph, weights=DEFAULT_WEIGHTS) -> float:
    w = _f32(weights)
    return float(lib().so_fuse_scores_full(_p(w, C.c_float), sem, ent, tag, imp, mom, int(access), graph))


def apply_recency_boost(score, now_s, created_s, boost_hours=24, multiplier=1.2) -> float:
    return float(lib().so_apply_recency_boost(score, int(now_s), int(created_s), boost_hours, multiplier))


def fuse_rank(feats: dict, now_s: int, max_results=5, min_score=0.25, weights=DEFAULT_WEIGHTS,
              recency_hours=24, recency_mult=1.2, graph_boost=1.15):
    """relevance.rs:805-918 ranking -> (order, scores, all_scores)."""
    w = _f32(weights)
    sem, ent, tag, imp, mom, gr = (_f32(feats[k]) for k in
                                   ("semantic", "entity", "tag", "importance", "momentum", "graph"))
    acc = np.ascontiguousarray(feats["access_count"], np.uint32)
    created = np.ascontiguousarray(feats["created_at"], np.int64)
    n = sem.size
    order = np.zeros(max(max_results, 1), np.uint32)
    score = np.zeros(max(max_results, 1), np.float32)
    allsc = np.zeros(max(n, 1), np.float32)
    r = lib().so_fuse_rank(_p(w, C.c_float), _p(sem, C.c_float), _p(ent, C.c_float),
                           _p(tag, C.c_float), _p(imp, C.c_float), _p(mom, C.c_float),
                           _p(acc, C.c_uint32), _p(gr, C.c_float), _p(created, C.c_int64), n,
                           int(now_s), recency_hours, recency_mult, graph_boost, min_score,
                           max_results, _p(order, C.c_uint32), _p(score, C.c_float),
                           _p(allsc, C.c_float))
    return order[:r].copy(), score[:r].copy(), allsc[:n].copy()


def pool_finalize(hidden_states, attention) -> np.ndarray:
    h = _f32(hidden_states)
    att = np.ascontiguousarray(attention, np.int64)
    seq, hidden = h.shape
    out = np.zeros(hidden, np.float32)
    lib().so_pool_finalize(_p(h, C.c_float), _p(att, C.c_int64), seq, hidden, _p(out, C.c_float))
    return out


def synth_corpus(seed: int, first_row: int, n: int, dim: int = 384, f16_round: bool = False) -> np.ndarray:
    out = np.empty((n, dim), np.float32)
    lib().so_synth_corpus(seed, first_row, n, dim, int(f16_round), _p(out, C.c_float))
    return out


def synth_queries(seed: int, n_rows: int, b: int, dim: int = 384):
    out = np.empty((b, dim), np.float32)
    near = np.empty(b, np.uint64)
    lib().so_synth_queries(seed, n_rows, b, dim, _p(out, C.c_float), _p(near, C.c_uint64))
    return out, near


def round_f16(a) -> np.ndarray:
    a = _f32(a)
    out = np.empty_like(a)
    lib().so_round_f16(_p(a, C.c_float), _p(out, C.c_float), a.size)
    return out


def max_threads() -> int:
    return int(lib().so_max_threads())


def rrf_fuse(k, weights, lists):
    """hybrid_search.rs::RRFusion::new(k, weights).fuse(lists).  lists: sequence of [n_i, 2] u64 arrays
    (MemoryId as uuid hi/lo) in rank order -> (ids [m,2] u64, scores [m] f32)."""
    weights = np.ascontiguousarray(weights, dtype=np.float32)
    lists = [np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 2) for x in lists]
    offsets = np.zeros(len(lists) + 1, dtype=np.uint32)
    offsets[1:] = np.cumsum([x.shape[0] for x in lists])
    ids = np.ascontiguousarray(np.concatenate(lists, axis=0) if lists else np.zeros((0, 2), np.uint64))
    n = int(offsets[-1])
    out_ids = np.zeros((max(n, 1), 2), dtype=np.uint64)
    out_sc = np.zeros(max(n, 1), dtype=np.float32)
    f = lib().so_rrf_fuse
    f.restype = C.c_uint32
    f.argtypes = [C.c_float, C.POINTER(C.c_float), C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_uint32,
                  C.POINTER(C.c_uint64), C.POINTER(C.c_float)]
    m = f(float(k), _p(weights, C.c_float), weights.size, _p(ids, C.c_uint64), _p(offsets, C.c_uint32), len(lists),
          _p(out_ids, C.c_uint64), _p(out_sc, C.c_float))
    return out_ids[:m].copy(), out_sc[:m].copy()


# ---------------------------------------------------------------------------------------------
# Vamana graph index (vamana_oracle.c): reference src/vector_db/vamana.rs
# ---------------------------------------------------------------------------------------------
class VamanaOracle:
    """VamanaIndex restated on the CPU: build / add_vector / search (ANN arm) / mark_deleted.

    Graph = deg [n] u32 + nbr [n, stride] u32 (stride = max_degree + 1), medoid.  Defaults are
    VamanaConfig::default (vamana.rs:80-91): R = 32, L = 75, alpha = 1.2, NormalizedDotProduct."""

    def __init__(self, dim, max_degree=32, search_list_size=75, alpha=1.2, metric=METRIC_NEG_DOT,
                 variant=VARIANT_SCALAR):
        self.dim, self.R, self.L, self.alpha = dim, max_degree, search_list_size, alpha
        self.metric, self.variant = metric, variant
        self.stride = max_degree + 1
        self.vectors = np.zeros((0, dim), np.float32)
        self.deg = np.zeros(0, np.uint32)
        self.nbr = np.zeros((0, self.stride), np.uint32)
        self.medoid = 0
        self.deleted = set()

    def __len__(self):
        return self.vectors.shape[0]

    def build(self, vectors, seed=0, init_graph=None):
        """VamanaIndex::build (vamana.rs:200-284); init_graph = (deg, nbr) overrides the seeded random start."""
        self.vectors = _f32(vectors).copy()
        n = len(self)
        self.deg = np.zeros(n, np.uint32)
        self.nbr = np.zeros((n, self.stride), np.uint32)
        if n == 0:
            return 0
        if init_graph is None:
            lib().so_vamana_init_graph(n, self.R, self.stride, seed, _p(self.deg, C.c_uint32), _p(self.nbr, C.c_uint32))
        else:
            d0, n0 = init_graph
            self.deg[:] = d0
            self.nbr[:, :n0.shape[1]] = n0
        self.medoid = int(lib().so_find_medoid(_p(self.vectors, C.c_float), n, self.dim, self.metric, self.variant))
        return int(lib().so_vamana_build(_p(self.vectors, C.c_float), n, self.dim, self.R, self.L, self.alpha,
                                         self.metric, self.variant, self.medoid, self.stride,
                                         _p(self.deg, C.c_uint32), _p(self.nbr, C.c_uint32)))

    def set_graph(self, vectors, deg, nbr, medoid):
        self.vectors = _f32(vectors).copy()
        n = len(self)
        self.deg = np.ascontiguousarray(deg, np.uint32).copy()
        self.nbr = np.zeros((n, self.stride), np.uint32)
        nbr = np.asarray(nbr, np.uint32)
        self.nbr[:, :nbr.shape[1]] = nbr[:, :self.stride]
        self.medoid = int(medoid)

    def add_vector(self, vector):
        """VamanaIndex::add_vector (vamana.rs:853-973) -> the new id."""
        n = len(self)
        self.vectors = np.ascontiguousarray(np.vstack([self.vectors, _f32(vector)[None, :]]))
        self.deg = np.concatenate([self.deg, np.zeros(1, np.uint32)])
        self.nbr = np.ascontiguousarray(np.vstack([self.nbr, np.zeros((1, self.stride), np.uint32)]))
        if n == 0:
            self.medoid = 0
        lib().so_vamana_add_vector(_p(self.vectors, C.c_float), n, self.dim, self.R, self.metric, self.variant,
                                   self.medoid, self.stride, _p(self.deg, C.c_uint32), _p(self.nbr, C.c_uint32))
        return n

    def mark_deleted(self, vid):
        self.deleted.add(int(vid))

    def greedy_search(self, query, k, entry=None, stats=False):
        query = _f32(query)
        ids = np.zeros(max(k, 1), np.uint32)
        dist = np.zeros(max(k, 1), np.float32)
        st = np.zeros(2, np.uint64)
        n = len(self)
        r = lib().so_greedy_search(_p(self.vectors, C.c_float), n, self.dim, _p(self.deg, C.c_uint32),
                                   _p(self.nbr, C.c_uint32), self.stride, n, _p(query, C.c_float), k,
                                   self.medoid if entry is None else entry, self.metric, self.variant,
                                   _p(ids, C.c_uint32), _p(dist, C.c_float), _p(st, C.c_uint64))
        return (ids[:r].copy(), dist[:r].copy(), st) if stats else (ids[:r].copy(), dist[:r].copy())

    def search(self, query, k, stats=False):
        """VamanaIndex::search, ANN arm (vamana.rs:779-807) -> (ids, dist)."""
        query = _f32(query)
        ids = np.zeros(max(k, 1), np.uint32)
        dist = np.zeros(max(k, 1), np.float32)
        st = np.zeros(2, np.uint64)
        n = len(self)
        bm = deleted_bitmap(n, self.deleted) if self.deleted else None
        r = lib().so_vamana_search(_p(self.vectors, C.c_float), n, self.dim, _p(self.deg, C.c_uint32),
                                   _p(self.nbr, C.c_uint32), self.stride, self.medoid,
                                   None if bm is None else _p(bm, C.c_uint32), len(self.deleted),
                                   _p(query, C.c_float), k, self.metric, self.variant, _p(ids, C.c_uint32),
                                   _p(dist, C.c_float), _p(st, C.c_uint64))
        return (ids[:r].copy(), dist[:r].copy(), st) if stats else (ids[:r].copy(), dist[:r].copy())

    def robust_prune(self, node_id, cand_ids, cand_dist):
        cand_ids = np.ascontiguousarray(cand_ids, np.uint32)
        cand_dist = _f32(cand_dist)
        out = np.zeros(self.R + 1, np.uint32)
        m = lib().so_robust_prune(_p(self.vectors, C.c_float), self.dim, node_id, _p(cand_ids, C.c_uint32),
                                  _p(cand_dist, C.c_float), cand_ids.size, self.R, self.alpha, self.metric,
                                  self.variant, _p(out, C.c_uint32))
        return out[:m].copy()


def find_medoid(corpus, metric=METRIC_NEG_DOT, variant=VARIANT_SCALAR) -> int:
    corpus = _f32(corpus)
    return int(lib().so_find_medoid(_p(corpus, C.c_float), corpus.shape[0], corpus.shape[1], metric, variant))


def recall_at_k(ann_ids, ann_counts, exact_ids, exact_counts) -> float:
    a = np.ascontiguousarray(ann_ids, np.uint32)
    e = np.ascontiguousarray(exact_ids, np.uint32)
    assert a.shape == e.shape
    ac = np.ascontiguousarray(ann_counts, np.uint32)
    ec = np.ascontiguousarray(exact_counts, np.uint32)
    return float(lib().so_recall_at_k(_p(a, C.c_uint32), _p(ac, C.c_uint32), _p(e, C.c_uint32), _p(ec, C.c_uint32),
                                      a.shape[0], a.shape[1]))


def clustered_corpus(seed, n, dim, n_clusters=None, spread=0.35):
    """Unit vectors around random centres (ANN needs structure; i.i.d. unit vectors in 384-d have none)."""
    rng = np.random.default_rng(seed)
    c = n_clusters or max(4, int(round(n ** 0.5 / 2)))
    centres = rng.standard_normal((c, dim)).astype(np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    which = rng.integers(0, c, n)
    x = centres[which] + (spread / np.sqrt(dim)) * rng.standard_normal((n, dim)).astype(np.float32) * np.sqrt(dim) / np.sqrt(dim)
    x = x.astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.ascontiguousarray(x, np.float32)


def manifold_corpus(seed, n, dim, latent=12, noise=0.05):
    """Unit vectors on a low-dimensional manifold: normalize(z A + noise), z ~ N(0, I_latent).  Sentence
    embeddings have low intrinsic dimension; the alpha-RNG prune of a Vamana graph relies on that (on tight
    isolated clusters or i.i.d. unit vectors every cluster-mate is "unoccluded" and the lists lose their
    long-range edges -- for the reference's build just the same)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((latent, dim)).astype(np.float32)
    out = np.empty((n, dim), np.float32)
    for lo in range(0, n, 100_000):
        m = min(100_000, n - lo)
        z = rng.standard_normal((m, latent)).astype(np.float32)
        x = z @ A + (noise * np.sqrt(latent)) * rng.standard_normal((m, dim)).astype(np.float32)
        out[lo:lo + m] = x / np.linalg.norm(x, axis=1, keepdims=True)
    return out


# ---------------------------------------------------------------------------------------------
# Product Quantization (pq_oracle.c): reference src/vector_db/pq.rs
# ---------------------------------------------------------------------------------------------
class PQOracle:
    """ProductQuantizer + CompressedVectorStore restated on the CPU.  centroids [m, k, ds] f32."""

    def __init__(self, dim=384, subvec_dim=8, num_centroids=256, kmeans_iterations=20):
        assert dim % subvec_dim == 0
        self.dim, self.ds, self.m = dim, subvec_dim, dim // subvec_dim
        self.num_centroids, self.iterations = num_centroids, kmeans_iterations
        self.centroids = None
        self.k = 0
        self.ids = np.zeros(0, np.uint32)
        self.codes = np.zeros((0, self.m), np.uint8)

    def shuffled_orders(self, n, seed=0):
        """one index permutation per subspace (the reference shuffles with thread_rng inside kmeans, pq.rs:157-159)"""
        rng = np.random.default_rng(seed)
        return np.stack([rng.permutation(n) for _ in range(self.m)]).astype(np.uint32)

    def train(self, vectors, seed=0, orders=None):
        v = _f32(vectors)
        n = v.shape[0]
        orders = self.shuffled_orders(n, seed) if orders is None else np.ascontiguousarray(orders, np.uint32)
        self.k = min(self.num_centroids, n)
        self.centroids = np.zeros((self.m, self.k, self.ds), np.float32)
        k = lib().so_pq_train(_p(v, C.c_float), n, self.dim, self.ds, self.num_centroids, self.iterations,
                              _p(orders, C.c_uint32), _p(self.centroids, C.c_float))
        assert k == self.k
        return self

    def encode(self, vectors):
        v = _f32(vectors).reshape(-1, self.dim)
        out = np.zeros((v.shape[0], self.m), np.uint8)
        lib().so_pq_encode(_p(self.centroids, C.c_float), self.m, self.k, self.ds, _p(v, C.c_float), v.shape[0],
                           _p(out, C.c_uint8))
        return out

    def decode(self, codes):
        c = np.ascontiguousarray(codes, np.uint8).reshape(-1, self.m)
        out = np.zeros((c.shape[0], self.dim), np.float32)
        lib().so_pq_decode(_p(self.centroids, C.c_float), self.m, self.k, self.ds, _p(c, C.c_uint8), c.shape[0],
                           _p(out, C.c_float))
        return out

    def table(self, query):
        q = _f32(query)
        t = np.zeros((self.m, self.k), np.float32)
        lib().so_pq_table(_p(self.centroids, C.c_float), self.m, self.k, self.ds, _p(q, C.c_float), _p(t, C.c_float))
        return t

    def distance_with_table(self, table, codes):
        t = _f32(table)
        c = np.ascontiguousarray(codes, np.uint8)
        return float(lib().so_pq_distance_with_table(_p(t, C.c_float), self.m, self.k, _p(c, C.c_uint8)))

    def asymmetric_distance(self, query, codes):
        q = _f32(query)
        c = np.ascontiguousarray(codes, np.uint8)
        return float(lib().so_pq_asymmetric_distance(_p(self.centroids, C.c_float), self.m, self.k, self.ds,
                                                     _p(q, C.c_float), _p(c, C.c_uint8)))

    def add(self, ids, vectors):
        """CompressedVectorStore::add for unique ids"""
        self.ids = np.concatenate([self.ids, np.ascontiguousarray(ids, np.uint32)])
        self.codes = np.ascontiguousarray(np.vstack([self.codes, self.encode(vectors)]))

    def search(self, query, topk):
        q = _f32(query)
        ids = np.zeros(max(topk, 1), np.uint32)
        dist = np.zeros(max(topk, 1), np.float32)
        r = lib().so_pq_search(_p(self.centroids, C.c_float), self.m, self.k, self.ds, _p(self.ids, C.c_uint32),
                               _p(self.codes, C.c_uint8), self.ids.size, _p(q, C.c_float), topk, _p(ids, C.c_uint32),
                               _p(dist, C.c_float))
        return ids[:r].copy(), dist[:r].copy()


# ---------------------------------------------------------------------------------------------
# SPANN partitioned index (spann_oracle.c): reference src/vector_db/spann.rs
# ---------------------------------------------------------------------------------------------
class SpannOracle:
    """SpannIndex restated on the CPU: build / insert / search.  Posting lists are kept per partition as
    (ids, codes) and flattened to CSR for the C search."""

    def __init__(self, dim=384, num_partitions=None, num_probes=10, kmeans_iterations=25, metric=METRIC_NEG_DOT,
                 use_pq=True, pq_iterations=20):
        self.dim, self.num_partitions_cfg, self.num_probes = dim, num_partitions, num_probes
        self.iterations, self.metric, self.use_pq, self.pq_iterations = kmeans_iterations, metric, use_pq, pq_iterations
        self.centroids = np.zeros((0, dim), np.float32)
        self.pq = None
        self.part_ids, self.part_codes = [], []

    def compute_partitions(self, n):
        """SpannConfig::compute_partitions (:136-139)"""
        import math
        return self.num_partitions_cfg or max(int(math.ceil(math.sqrt(float(n)))), 1)

    def __len__(self):
        return sum(len(p) for p in self.part_ids)

    def build(self, vectors, seed=0, order=None, pq_orders=None):
        v = _f32(vectors)
        if v.shape[0] == 0:
            raise ValueError("Cannot build index from empty vectors")
        if not self.use_pq:
            raise ValueError("SPANN requires PQ (use_pq=true): posting lists store only PQ codes")
        if v.shape[1] != self.dim:
            raise ValueError(f"Vector dimension {v.shape[1]} doesn't match config {self.dim}")
        n = v.shape[0]
        k = self.compute_partitions(n)
        rng = np.random.default_rng(seed)
        self.order = rng.permutation(n).astype(np.uint32) if order is None else np.ascontiguousarray(order, np.uint32)
        self.centroids = np.zeros((k, self.dim), np.float32)
        self.kmeans_iterations_run = int(lib().so_spann_kmeans(_p(v, C.c_float), n, self.dim, k, self.iterations, self.metric,
                                                               _p(self.order, C.c_uint32), _p(self.centroids, C.c_float)))
        self.pq = PQOracle(self.dim, kmeans_iterations=self.pq_iterations)
        self.pq_orders = self.pq.shuffled_orders(n, seed + 1) if pq_orders is None else pq_orders
        self.pq.train(v, orders=self.pq_orders)
        part = np.zeros(n, np.uint32)
        lib().so_spann_assign(_p(v, C.c_float), n, self.dim, _p(self.centroids, C.c_float), k, self.metric, _p(part, C.c_uint32))
        codes = self.pq.encode(v)
        self.part_ids = [np.nonzero(part == p)[0].astype(np.uint32) for p in range(k)]       # vector order inside a partition
        self.part_codes = [codes[i] for i in self.part_ids]
        return self

    def insert(self, vector_id, vector):
        """SpannIndex::insert (:1006-1051)"""
        if self.centroids.shape[0] == 0:
            raise ValueError("Cannot insert into empty index - build first")
        v = _f32(vector)
        p = int(lib().so_spann_nearest_centroid(_p(v, C.c_float), _p(self.centroids, C.c_float), self.centroids.shape[0],
                                                self.dim, self.metric))
        self.part_ids[p] = np.concatenate([self.part_ids[p], np.array([vector_id], np.uint32)])
        self.part_codes[p] = np.vstack([self.part_codes[p], self.pq.encode(v[None, :])])
        return p

    def csr(self):
        off = np.zeros(len(self.part_ids) + 1, np.uint64)
        off[1:] = np.cumsum([len(p) for p in self.part_ids])
        ids = np.concatenate(self.part_ids).astype(np.uint32) if len(self) else np.zeros(0, np.uint32)
        codes = np.ascontiguousarray(np.vstack(self.part_codes), np.uint8) if len(self) else np.zeros((0, self.pq.m), np.uint8)
        return off, ids, codes

    def search(self, query, k, with_probes=False):
        q = _f32(query)
        if self.centroids.shape[0] == 0:
            return (np.zeros(0, np.uint32), np.zeros(0, np.float32))
        if q.size != self.dim:
            raise ValueError(f"Query dimension {q.size} doesn't match index dimension {self.dim}")
        off, ids, codes = self.csr()
        out_i = np.zeros(max(k, 1), np.uint32)
        out_d = np.zeros(max(k, 1), np.float32)
        probes = np.zeros(max(self.num_probes, 1), np.uint32)
        r = lib().so_spann_search(_p(self.centroids, C.c_float), self.centroids.shape[0], self.dim, self.metric, self.num_probes,
                                  _p(off, C.c_uint64), _p(ids, C.c_uint32), _p(codes, C.c_uint8),
                                  _p(self.pq.centroids, C.c_float), self.pq.m, self.pq.k, self.pq.ds, _p(q, C.c_float), k,
                                  _p(out_i, C.c_uint32), _p(out_d, C.c_float), _p(probes, C.c_uint32))
        res = (out_i[:r].copy(), out_d[:r].copy())
        return res + (probes[:min(self.num_probes, self.centroids.shape[0])].copy(),) if with_probes else res
