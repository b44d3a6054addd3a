"""ctypes binding of libshodh_gpu.so (the C ABI declared in include/shodh_gpu.h).

Loading fails loudly if the CUDA library is missing -- there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libshodh_gpu.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "shodh_gpu.h")

SG_OK = 0
SG_ERR_INVALID, SG_ERR_CUDA, SG_ERR_OOM, SG_ERR_DIM = -1, -2, -3, -4
SG_ERR_UNSUPPORTED, SG_ERR_NO_DEVICE, SG_ERR_NOT_BUILT = -5, -6, -7
SG_DTYPE_F32, SG_DTYPE_F16 = 0, 1
SG_METRIC_NEG_DOT, SG_METRIC_EUCLIDEAN, SG_METRIC_COSINE, SG_METRIC_COSINE_SIM = 0, 1, 2, 3
SG_ORDER_SCALAR, SG_ORDER_AVX2 = 0, 1


class SgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libshodh_gpu error {code}: {msg}")
        self.code = code


class FuseWeights(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("semantic", "entity", "tag", "importance", "momentum",
                                         "access_count", "graph_strength")]


class FuseConfig(C.Structure):
    _fields_ = [("recency_boost_hours", C.c_uint64), ("recency_boost_multiplier", C.c_float),
                ("graph_boost_multiplier", C.c_float), ("min_score", C.c_float),
                ("max_results", C.c_uint32)]


class CandidateFeats(C.Structure):
    _fields_ = [("semantic", C.POINTER(C.c_float)), ("entity", C.POINTER(C.c_float)),
                ("tag", C.POINTER(C.c_float)), ("importance", C.POINTER(C.c_float)),
                ("momentum", C.POINTER(C.c_float)), ("access_count", C.POINTER(C.c_uint32)),
                ("graph_strength", C.POINTER(C.c_float)), ("created_at_unix_s", C.POINTER(C.c_int64))]


def declared_symbols() -> list[str]:
    """Every function include/shodh_gpu.h declares with SG_API."""
    src = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"SG_API\s+[\w\s\*]+?\b(sg_\w+)\s*\(", src)))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "shodh_memory_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    if missing:
        raise ImportError(f"libshodh_gpu.so does not export: {missing}")
    vp, u32p, f32p, u64p = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.POINTER(C.c_uint64)
    L.sg_version.restype = C.c_char_p
    L.sg_last_error.restype = C.c_char_p
    L.sg_kernel_launch_count.restype = C.c_uint64
    L.sg_device_count.argtypes = [C.POINTER(C.c_int)]
    L.sg_store_create.argtypes = [C.POINTER(vp), C.c_int, C.c_uint32, C.c_int, C.c_int]
    L.sg_store_destroy.argtypes = [vp]
    L.sg_store_destroy.restype = None
    L.sg_store_reserve.argtypes = [vp, C.c_uint64]
    L.sg_store_append.argtypes = [vp, f32p, C.c_uint64, u32p]
    L.sg_store_append_device.argtypes = [vp, vp, C.c_uint64, u32p, vp]
    L.sg_store_append_synthetic.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64]
    L.sg_store_mark_deleted.argtypes = [vp, C.c_uint32, C.POINTER(C.c_int)]
    L.sg_store_is_deleted.argtypes = [vp, C.c_uint32, C.POINTER(C.c_int)]
    L.sg_store_clear_deleted.argtypes = [vp]
    L.sg_store_len.argtypes = [vp, u64p]
    L.sg_store_memory_info.argtypes = [vp, C.POINTER(C.c_int), u64p, u64p, u32p]
    L.sg_store_deleted_count.argtypes = [vp, u64p]
    L.sg_store_dim.argtypes = [vp, u32p]
    L.sg_store_get_vectors.argtypes = [vp, C.c_uint64, C.c_uint64, f32p]
    L.sg_store_load_vama.argtypes = [vp, C.c_char_p, C.c_int]
    L.sg_store_save_vama.argtypes = [vp, C.c_char_p]
    L.sg_vama_inspect.argtypes = [C.c_char_p, C.c_int, u64p, u32p, u32p, u32p, u64p]
    L.sg_store_set_order.argtypes = [vp, C.c_int]
    L.sg_store_set_id_base.argtypes = [vp, C.c_uint32]
    L.sg_store_set_profiling.argtypes = [vp, C.c_int]
    L.sg_store_last_scan_ms.argtypes = [vp, f32p, u32p]
    L.sg_store_search.argtypes = [vp, f32p, C.c_uint32, C.c_uint32, u32p, f32p, u32p]
    L.sg_top_k_similar.argtypes = [C.c_int, f32p, C.c_uint32, f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, u32p, f32p,
                                   u32p]
    L.sg_store_set_memory_ids.argtypes = [vp, u64p, C.c_uint64]
    L.sg_store_search_memory_ids.argtypes = [vp, f32p, C.c_uint32, C.c_uint32, u64p, f32p, u32p]
    L.sg_store_search_device.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp]
    L.sg_merge_topk_device.argtypes = [C.c_int, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, vp, vp]
    L.sg_merge_topk_device_strided.argtypes = [C.c_int, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                               C.c_uint64, vp, vp, vp, vp]
    L.sg_exchange_buffer_words.restype = C.c_uint64
    L.sg_exchange_buffer_words.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    L.sg_exchange_merge_topk_device.argtypes = [C.c_int, u64p, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32,
                                                C.c_uint32, vp, vp, vp, vp]
    L.sg_store_tc_stats.argtypes = [vp, u64p, u64p]
    L.sg_store_set_tc_mode.argtypes = [vp, C.c_int]
    L.sg_debug_tc_scores.argtypes = [vp, f32p, C.c_uint32, C.c_uint64, C.c_uint64, f32p]
    L.sg_debug_exact_scores.argtypes = [vp, f32p, C.c_uint32, C.c_uint64, C.c_uint64, f32p]
    L.sg_store_search_sharded_device.argtypes = [vp, vp, C.c_uint32, C.c_uint32, u64p, C.c_uint32, C.c_uint32, C.c_uint32,
                                                 vp, vp, vp, vp]
    L.sg_store_metric.argtypes = [vp, C.POINTER(C.c_int)]
    L.sg_store_enable_graph.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_float]
    L.sg_store_build_graph.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_float, C.c_uint64]
    L.sg_store_set_graph.argtypes = [vp, C.c_uint32, C.c_uint32, u32p, u32p, C.c_uint32, C.c_uint64]
    L.sg_store_get_graph.argtypes = [vp, u32p, u32p, C.c_uint64]
    L.sg_store_graph_info.argtypes = [vp, u32p, u32p, u64p]
    L.sg_store_search_ann.argtypes = [vp, f32p, C.c_uint32, C.c_uint32, u32p, f32p, u32p, u64p]
    L.sg_store_search_ann_device.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp]
    L.sg_store_ann_stats.argtypes = [vp, u64p, u64p]
    L.sg_store_ann_phase_cycles.argtypes = [vp, C.c_int, u64p]
    L.sg_store_profile_read.argtypes = [vp, f32p, C.c_uint32, u32p, u32p]
    i64p = C.POINTER(C.c_int64)
    u8p = C.POINTER(C.c_uint8)
    L.sg_pq_create.argtypes = [C.POINTER(vp), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, f32p]
    L.sg_pq_train.argtypes = [C.POINTER(vp), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, f32p, C.c_uint64, u32p]
    L.sg_pq_destroy.argtypes = [vp]
    L.sg_pq_destroy.restype = None
    L.sg_pq_info.argtypes = [vp, u32p, u32p, u32p, u64p]
    L.sg_pq_get_centroids.argtypes = [vp, f32p]
    L.sg_pq_encode.argtypes = [vp, f32p, C.c_uint64, u8p]
    L.sg_pq_decode.argtypes = [vp, u8p, C.c_uint64, f32p]
    L.sg_pq_table.argtypes = [vp, f32p, f32p]
    L.sg_pq_add.argtypes = [vp, u32p, f32p, C.c_uint64]
    L.sg_pq_search.argtypes = [vp, f32p, C.c_uint32, C.c_uint32, u32p, f32p, u32p]
    L.sg_spann_build.argtypes = [C.POINTER(vp), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, f32p, C.c_uint64,
                                 u32p, C.c_uint32, u32p]
    L.sg_spann_create.argtypes = [C.POINTER(vp), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, f32p, f32p, C.c_uint32,
                                  C.c_uint32, u64p, u32p, u8p]
    L.sg_spann_destroy.argtypes = [vp]
    L.sg_spann_destroy.restype = None
    L.sg_spann_info.argtypes = [vp, u32p, u64p, u32p]
    L.sg_spann_get.argtypes = [vp, f32p, f32p, u64p, u32p, u8p]
    L.sg_spann_set_probes.argtypes = [vp, C.c_uint32]
    L.sg_spann_shape.argtypes = [vp, u32p, u32p, u32p, u32p, C.POINTER(C.c_int)]
    L.sg_spann_file_inspect.argtypes = [C.c_char_p, C.c_int, u64p, u32p, u32p, u32p, u32p, u32p, u32p]
    L.sg_spann_file_read.argtypes = [C.c_char_p, f32p, f32p, u64p, u32p, u8p]
    L.sg_spann_file_write.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, f32p, C.c_uint32, C.c_uint32, C.c_uint32,
                                      f32p, u64p, u32p, u8p]
    L.sg_spann_load_file.argtypes = [C.POINTER(vp), C.c_int, C.c_char_p, C.c_uint32, C.c_int]
    L.sg_spann_save_file.argtypes = [vp, C.c_char_p]
    L.sg_spann_insert.argtypes = [vp, C.c_uint32, f32p]
    L.sg_spann_search.argtypes = [vp, f32p, C.c_uint32, C.c_uint32, u32p, f32p, u32p]
    L.sg_encoder_create.argtypes = [C.POINTER(vp), C.c_int, C.c_char_p]
    L.sg_encoder_destroy.argtypes = [vp]
    L.sg_encoder_destroy.restype = None
    L.sg_encoder_dim.argtypes = [vp, u32p]
    L.sg_encoder_encode.argtypes = [vp, i64p, i64p, C.c_uint32, C.c_uint32, f32p]
    L.sg_encoder_encode_device.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, vp, vp]
    L.sg_pool_finalize.argtypes = [C.c_int, f32p, i64p, C.c_uint32, C.c_uint32, C.c_uint32, f32p]
    L.sg_fuse_default_weights.argtypes = [C.POINTER(FuseWeights)]
    L.sg_fuse_default_weights.restype = None
    L.sg_fuse_default_config.argtypes = [C.POINTER(FuseConfig)]
    L.sg_fuse_default_config.restype = None
    L.sg_fuse_topk.argtypes = [C.c_int, C.POINTER(FuseWeights), C.POINTER(FuseConfig),
                               C.POINTER(CandidateFeats), C.c_uint32, C.c_int64, u32p, f32p, u32p, f32p]
    L.sg_rrf_fuse.argtypes = [C.c_int, C.c_float, f32p, C.c_uint32, C.POINTER(C.c_uint64), u32p, C.c_uint32,
                              C.POINTER(C.c_uint64), f32p, u32p]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != SG_OK:
        raise SgError(rc, lib().sg_last_error().decode(errors="replace"))
