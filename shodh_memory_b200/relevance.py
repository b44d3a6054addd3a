"""Host-side mirror of the ranking phase of RelevanceEngine::surface_relevant_inner
(reference src/relevance.rs:805-918) over sg_fuse_topk."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def default_weights() -> capi.FuseWeights:
    w = capi.FuseWeights()
    capi.lib().sg_fuse_default_weights(C.byref(w))
    return w


def default_config() -> capi.FuseConfig:
    c = capi.FuseConfig()
    capi.lib().sg_fuse_default_config(C.byref(c))
    return c


def fuse_topk(feats: dict, now_unix_s: int, max_results: int = 5, min_score: float = 0.25,
              weights=None, device: int = 0, recency_hours: int = 24, recency_mult: float = 1.2,
              graph_boost: float = 1.15):
    """feats: semantic/entity/tag/importance/momentum/graph f32[n], access_count u32[n],
    created_at i64[n] -> (order u32[m], scores f32[m], all_scores f32[n])."""
    w = default_weights()
    if weights is not None:
        for name, v in zip(("semantic", "entity", "tag", "importance", "momentum", "access_count",
                            "graph_strength"), weights):
            setattr(w, name, float(v))
    cfg = default_config()
    cfg.max_results, cfg.min_score = max_results, min_score
    cfg.recency_boost_hours, cfg.recency_boost_multiplier = recency_hours, recency_mult
    cfg.graph_boost_multiplier = graph_boost
    arr = {k: np.ascontiguousarray(feats[k], np.float32) for k in
           ("semantic", "entity", "tag", "importance", "momentum", "graph")}
    acc = np.ascontiguousarray(feats["access_count"], np.uint32)
    created = np.ascontiguousarray(feats["created_at"], np.int64)
    n = acc.size
    f = capi.CandidateFeats()
    f.semantic = arr["semantic"].ctypes.data_as(C.POINTER(C.c_float))
    f.entity = arr["entity"].ctypes.data_as(C.POINTER(C.c_float))
    f.tag = arr["tag"].ctypes.data_as(C.POINTER(C.c_float))
    f.importance = arr["importance"].ctypes.data_as(C.POINTER(C.c_float))
    f.momentum = arr["momentum"].ctypes.data_as(C.POINTER(C.c_float))
    f.graph_strength = arr["graph"].ctypes.data_as(C.POINTER(C.c_float))
    f.access_count = acc.ctypes.data_as(C.POINTER(C.c_uint32))
    f.created_at_unix_s = created.ctypes.data_as(C.POINTER(C.c_int64))
    order = np.zeros(max(max_results, 1), np.uint32)
    score = np.zeros(max(max_results, 1), np.float32)
    allsc = np.zeros(max(n, 1), np.float32)
    m = C.c_uint32()
    capi.check(capi.lib().sg_fuse_topk(device, C.byref(w), C.byref(cfg), C.byref(f), n, int(now_unix_s),
                                       order.ctypes.data_as(C.POINTER(C.c_uint32)),
                                       score.ctypes.data_as(C.POINTER(C.c_float)), C.byref(m),
                                       allsc.ctypes.data_as(C.POINTER(C.c_float))))
    return order[:m.value].copy(), score[:m.value].copy(), allsc[:n].copy()


class RRFusion:
    """Mirror of `RRFusion` (reference src/memory/hybrid_search.rs:529-594): `RRFusion(k, weights).fuse(lists)`.

    A ranked list is an `[n, 2]` u64 array of MemoryIds (Uuid as two big-endian halves) in rank order -- what
    `GpuVectorStore.search_ids` returns for the vector leg.  Returns (ids [m,2] u64, rrf_scores [m] f32) sorted
    (score desc by total_cmp, MemoryId asc).  Weights are passed RAW; the library normalises them like
    `RRFusion::new`."""

    def __init__(self, k: float, weights, device: int = 0):
        self.k = float(k)
        self.weights = np.ascontiguousarray(weights, dtype=np.float32)
        self.device = int(device)

    def fuse(self, ranked_lists):
        lists = [np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 2) for x in ranked_lists]
        offsets = np.zeros(len(lists) + 1, dtype=np.uint32)
        if lists:
            offsets[1:] = np.cumsum([x.shape[0] for x in lists])
        n = int(offsets[-1])
        ids = np.ascontiguousarray(np.concatenate(lists, axis=0)) if n else np.zeros((1, 2), np.uint64)
        out_ids = np.zeros((max(n, 1), 2), dtype=np.uint64)
        out_sc = np.zeros(max(n, 1), dtype=np.float32)
        m = C.c_uint32(0)
        u64p = C.POINTER(C.c_uint64)
        capi.check(capi.lib().sg_rrf_fuse(self.device, self.k, self.weights.ctypes.data_as(C.POINTER(C.c_float)),
                                          self.weights.size, ids.ctypes.data_as(u64p),
                                          offsets.ctypes.data_as(C.POINTER(C.c_uint32)), len(lists),
                                          out_ids.ctypes.data_as(u64p), out_sc.ctypes.data_as(C.POINTER(C.c_float)),
                                          C.byref(m)))
        return out_ids[:m.value].copy(), out_sc[:m.value].copy()
