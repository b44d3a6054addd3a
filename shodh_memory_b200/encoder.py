"""Host-side mirror of the reference's `Embedder` surface (src/embeddings/mod.rs:52-88) over the GPU
encoder.  Tokenisation is the host's job (the reference uses the `tokenizers` crate); this class takes
token ids + attention masks, exactly the tensors the reference feeds its ONNX session
(src/embeddings/minilm.rs:924-934)."""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from . import capi


def export_bert_weights(state_dict, path: str, layers: int, hidden: int, heads: int, ffn: int, vocab: int,
                        max_pos: int) -> None:
    """Write a HuggingFace BertModel state_dict (tensors or numpy arrays) as an SGW1 blob."""
    items = []
    for name, t in state_dict.items():
        if name.startswith("bert."):
            name = name[5:]
        if name.startswith("pooler.") or name.endswith("position_ids"):
            continue
        a = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
        items.append((name, np.ascontiguousarray(a, np.float32)))
    with open(path, "wb") as f:
        f.write(b"SGW1")
        f.write(struct.pack("<I6II", 1, layers, hidden, heads, ffn, vocab, max_pos, len(items)))
        for name, a in items:
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)) + nb)
            f.write(struct.pack("<I", a.ndim) + struct.pack(f"<{a.ndim}I", *a.shape))
            f.write(a.tobytes())


def random_minilm_state_dict(seed: int = 0, layers: int = 6, hidden: int = 384, ffn: int = 1536, vocab: int = 30522,
                             max_pos: int = 512) -> dict:
    """Seeded random weights with the MiniLM-L6 parameter names/shapes (no checkpoint is available offline)."""
    rng = np.random.default_rng(seed)
    def w(*shape, s=0.05):
        return (rng.standard_normal(shape) * s).astype(np.float32)
    sd = {"embeddings.word_embeddings.weight": w(vocab, hidden), "embeddings.position_embeddings.weight": w(max_pos, hidden),
          "embeddings.token_type_embeddings.weight": w(2, hidden),
          "embeddings.LayerNorm.weight": 1 + w(hidden, s=0.1), "embeddings.LayerNorm.bias": w(hidden)}
    for l in range(layers):
        p = f"encoder.layer.{l}."
        for nm in ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense"):
            sd[p + nm + ".weight"], sd[p + nm + ".bias"] = w(hidden, hidden, s=0.06), w(hidden)
        sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"] = w(ffn, hidden, s=0.06), w(ffn)
        sd[p + "output.dense.weight"], sd[p + "output.dense.bias"] = w(hidden, ffn, s=0.06), w(hidden)
        for nm in ("attention.output.LayerNorm", "output.LayerNorm"):
            sd[p + nm + ".weight"], sd[p + nm + ".bias"] = 1 + w(hidden, s=0.1), w(hidden)
    return sd


def pool_finalize(hidden_states, attention_mask, device: int = 0) -> np.ndarray:
    """The tail of generate_embedding_onnx on the device (minilm.rs:959-978 masked mean-pool + :846-878 finalize_pooled):
    hidden_states [b,L,H] f32, attention_mask [b,L] -> [b,H], bit-identical to the CPU path (sg_pool_finalize)."""
    h = np.ascontiguousarray(hidden_states, np.float32)
    m = np.ascontiguousarray(attention_mask, np.int64)
    if h.ndim != 3 or m.shape != h.shape[:2]:
        raise capi.SgError(capi.SG_ERR_DIM, f"expected [b,L,H] hidden states and a [b,L] mask, got {h.shape} / {m.shape}")
    b, L, H = h.shape
    out = np.empty((b, H), np.float32)
    capi.check(capi.lib().sg_pool_finalize(device, h.ctypes.data_as(C.POINTER(C.c_float)),
                                           m.ctypes.data_as(C.POINTER(C.c_int64)), b, L, H,
                                           out.ctypes.data_as(C.POINTER(C.c_float))))
    return out


class GpuMiniLM:
    """`impl Embedder` on token ids: dimension(), encode_ids() (= encode / encode_batch after tokenising)."""

    def __init__(self, weights_path: str, device: int = 0):
        self._h = C.c_void_p()
        capi.check(capi.lib().sg_encoder_create(C.byref(self._h), device, weights_path.encode()))
        d = C.c_uint32()
        capi.check(capi.lib().sg_encoder_dim(self._h, C.byref(d)))
        self._dim = d.value

    def dimension(self) -> int:
        return self._dim

    def encode_ids(self, input_ids, attention_mask) -> np.ndarray:
        ids = np.ascontiguousarray(input_ids, np.int64)
        mask = np.ascontiguousarray(attention_mask, np.int64)
        if ids.ndim != 2 or ids.shape != mask.shape:
            raise capi.SgError(capi.SG_ERR_DIM, f"expected [b,L] ids and mask, got {ids.shape} / {mask.shape}")
        b, L = ids.shape
        out = np.empty((b, self._dim), np.float32)
        capi.check(capi.lib().sg_encoder_encode(self._h, ids.ctypes.data_as(C.POINTER(C.c_int64)),
                                                mask.ctypes.data_as(C.POINTER(C.c_int64)), b, L,
                                                out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def encode_ids_device(self, d_ids: int, d_mask: int, b: int, L: int, d_out: int, stream: int = 0) -> None:
        capi.check(capi.lib().sg_encoder_encode_device(self._h, d_ids, d_mask, b, L, d_out, stream))

    def close(self) -> None:
        if self._h:
            capi.lib().sg_encoder_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
