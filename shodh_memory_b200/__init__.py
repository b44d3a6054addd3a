"""shodh_memory_b200 -- B200-native (sm_100a) semantic-recall hot path of shodh-memory.

Only what the path needs: csrc/ (CUDA kernels + the C ABI in include/shodh_gpu.h), and a thin
host-side mirror of the reference's VamanaIndex / relevance surface for tests and bench.
Importing the package loads lib/libshodh_gpu.so and fails loudly if it is missing."""
from . import capi
from .capi import SgError
from .encoder import GpuMiniLM, export_bert_weights, pool_finalize, random_minilm_state_dict
from .pq import GpuProductQuantizer
from .relevance import fuse_topk, RRFusion
from .spann import GpuSpannIndex, spann_file_inspect, spann_file_read, spann_file_write
from .vector_store import GpuVectorStore, merge_topk_device, vama_inspect, top_k_similar, cosine_similarity

capi.lib()  # load now: no lazy CPU fallback, a missing extension is an ImportError


def version() -> str:
    return capi.lib().sg_version().decode()


def kernel_launch_count() -> int:
    return int(capi.lib().sg_kernel_launch_count())


def device_count() -> int:
    import ctypes as C
    n = C.c_int()
    rc = capi.lib().sg_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def __getattr__(name):  # torch-dependent helpers load lazily (the C ABI itself needs no torch)
    if name in ("ShardedGpuStore", "shard_range", "gathered_strides", "exchange"):
        from . import sharded
        return getattr(sharded, name)
    raise AttributeError(name)


__all__ = ["GpuVectorStore", "ShardedGpuStore", "GpuMiniLM", "GpuProductQuantizer", "GpuSpannIndex", "export_bert_weights", "merge_topk_device", "top_k_similar", "cosine_similarity", "fuse_topk", "RRFusion", "SgError", "version", "kernel_launch_count",
           "device_count", "capi"]
