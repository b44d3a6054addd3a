"""GPU encoder vs a PyTorch fp32 reference of the same op.

Parity with the reference itself is UNPINNED for embedding values (the reference runs a third-party
ONNX Runtime graph on weights it downloads at first run; neither is in this environment).  The
checker is therefore a HuggingFace `BertModel` (fp32) with the MiniLM-L6 architecture and seeded
random weights, followed by the oracle's restatement of the reference pooling (minilm.rs:959-978,
:846-878).  Tolerance (fp16 tensor-core GEMMs, fp16 activations between layers, f32 accumulation):
cosine >= 0.999 and max |diff| <= 6e-3 on the unit-norm 384-d embedding -- NOT the 1e-4 of the scan."""
import os

import numpy as np
import pytest


pytestmark = pytest.mark.gpu
COS_TOL, ABS_TOL = 0.999, 6e-3
CFG = dict(layers=6, hidden=384, heads=12, ffn=1536, vocab=30522, max_pos=512)


@pytest.fixture(scope="module")
def model_and_encoder(sg, tmp_path_factory):
    import torch
    from transformers import BertConfig, BertModel
    torch.manual_seed(1234)
    cfg = BertConfig(vocab_size=CFG["vocab"], hidden_size=CFG["hidden"], num_hidden_layers=CFG["layers"],
                     num_attention_heads=CFG["heads"], intermediate_size=CFG["ffn"], max_position_embeddings=CFG["max_pos"],
                     type_vocab_size=2, layer_norm_eps=1e-12, hidden_act="gelu", hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0)
    model = BertModel(cfg, add_pooling_layer=False).eval()
    with torch.no_grad():                       # make biases / LayerNorm params non-trivial, weights MiniLM-sized
        for name, p in model.named_parameters():
            if name.endswith("LayerNorm.weight"):
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
            elif name.endswith("bias"):
                p.copy_(0.05 * torch.randn_like(p))
            elif "embeddings" in name:
                p.copy_(0.05 * torch.randn_like(p))
            else:
                p.copy_(0.06 * torch.randn_like(p))
    path = str(tmp_path_factory.mktemp("w") / "minilm_random.sgw")
    sg.export_bert_weights(model.state_dict(), path, **CFG)
    enc = sg.GpuMiniLM(path)
    yield model, enc
    enc.close()


def reference_embed(model, oracle, ids, mask):
    import torch
    with torch.no_grad():
        h = model(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask)).last_hidden_state.numpy()
    return np.stack([oracle.pool_finalize(h[i], mask[i]) for i in range(ids.shape[0])])


def make_batch(b, L, seed, ragged=True):
    rng = np.random.default_rng(seed)
    ids = rng.integers(1000, CFG["vocab"], (b, L)).astype(np.int64)
    ids[:, 0] = 101
    mask = np.ones((b, L), np.int64)
    if ragged:
        for i in range(b):
            n = int(rng.integers(2, L + 1))
            ids[i, n - 1] = 102
            ids[i, n:] = 0
            mask[i, n:] = 0
    else:
        ids[:, -1] = 102
    return ids, mask


@pytest.mark.parametrize("b,L", [(1, 64), (3, 17), (8, 128), (5, 256), (33, 64), (160, 64), (320, 64)])
def test_encoder_matches_torch_fp32(model_and_encoder, oracle, b, L):
    model, enc = model_and_encoder
    assert enc.dimension() == 384                                    # minilm.rs:1206-1208
    ids, mask = make_batch(b, L, seed=b * 1000 + L)
    got = enc.encode_ids(ids, mask)
    want = reference_embed(model, oracle, ids, mask)
    assert got.shape == (b, 384) and np.isfinite(got).all()
    assert np.allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)   # minilm.rs:1416-1417
    cos = (got * want).sum(1)
    assert cos.min() >= COS_TOL, cos
    assert np.abs(got - want).max() <= ABS_TOL, np.abs(got - want).max()


def test_config5_query_and_padding_invariance(model_and_encoder, oracle):
    """BASELINE config 5 shape (one 64-token query).  The fp32 graph is pad-invariant
    (minilm.rs:593-594): padding the same tokens to 256 must not change the embedding."""
    model, enc = model_and_encoder
    ids, mask = make_batch(1, 64, seed=7, ragged=False)
    e64 = enc.encode_ids(ids, mask)
    ids256 = np.zeros((1, 256), np.int64); ids256[:, :64] = ids
    mask256 = np.zeros((1, 256), np.int64); mask256[:, :64] = 1
    e256 = enc.encode_ids(ids256, mask256)
    assert float((e64 * e256).sum()) >= 0.99999
    want = reference_embed(model, oracle, ids, mask)
    assert float((e64 * want).sum()) >= COS_TOL


def test_all_masked_row_gives_zeros(model_and_encoder):
    """Empty text -> 384 zeros (minilm.rs:1123-1125); an all-masked row must not poison its batch."""
    model, enc = model_and_encoder
    ids, mask = make_batch(4, 32, seed=3)
    mask[2, :] = 0
    out = enc.encode_ids(ids, mask)
    assert not out[2].any()
    assert np.allclose(np.linalg.norm(out[[0, 1, 3]], axis=1), 1.0, atol=1e-5)


def test_encoder_feeds_the_store_on_device(model_and_encoder, sg, oracle):
    """config 5 end to end on the GPU: ids -> embedding (device) -> exact scan -> top-10."""
    import torch
    model, enc = model_and_encoder
    n = 50_000
    st = sg.GpuVectorStore(384, "f16")
    st.append_synthetic(23821, 0, n)
    ids, mask = make_batch(1, 64, seed=11, ragged=False)
    d_ids, d_mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    emb = torch.empty((1, 384), dtype=torch.float32, device="cuda")
    out_i = torch.empty((1, 10), dtype=torch.int32, device="cuda")
    out_d = torch.empty((1, 10), dtype=torch.float32, device="cuda")
    out_c = torch.empty(1, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    enc.encode_ids_device(d_ids.data_ptr(), d_mask.data_ptr(), 1, 64, emb.data_ptr(), stream)
    st.search_device(emb.data_ptr(), 1, 10, out_i.data_ptr(), out_d.data_ptr(), out_c.data_ptr(), stream)
    torch.cuda.synchronize()
    q = emb.cpu().numpy()
    seen = oracle.synth_corpus(23821, 0, n, 384, f16_round=True)
    rid, rdist = oracle.brute_force_search(seen, q[0], 10)       # scan parity given the GPU's own embedding
    assert np.array_equal(out_i.cpu().numpy().view(np.uint32)[0], rid)
    assert np.array_equal(out_d.cpu().numpy().view(np.uint32)[0], rdist.view(np.uint32))
    st.close()


def test_batched_document_ingest_stays_on_the_device(model_and_encoder, sg):
    """SURVEY 8f-4 batched ingest (retrieval.rs:668-676 -> minilm.rs:996): encode_batch output goes straight into
    the store (sg_encoder_encode_device -> sg_store_append_device), no host round trip for the embeddings."""
    import torch
    model, enc = model_and_encoder
    ids, mask = make_batch(48, 64, seed=21)
    want = enc.encode_ids(ids, mask)                                   # the host-buffer form of the same encoder
    d_ids, d_mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    emb = torch.empty((48, 384), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    st = sg.GpuVectorStore(384, "f32")
    enc.encode_ids_device(d_ids.data_ptr(), d_mask.data_ptr(), 48, 64, emb.data_ptr(), stream)
    assert st.add_vectors_device(emb.data_ptr(), 48, stream) == 0      # same stream: ordered after the encoder
    assert len(st) == 48
    got = st.extract_all_vectors()
    assert np.abs(got - want).max() <= 1e-6
    for i in (0, 17, 47):                                              # every document finds itself
        assert st.search(got[i], 1)[0][0] == i
    st.close()


def test_pool_finalize_is_bit_identical_incl_scrub_and_epsilon_branches(sg, oracle):
    """a3/a4 in isolation: f32 hidden states -> pool_norm_kernel vs oracle so_pool_finalize
    (minilm.rs:959-978, :846-878), bit for bit, including the NaN/Inf scrub and the `norm > f32::EPSILON` divide."""
    rng = np.random.default_rng(7)
    b, L, H = 9, 23, 384
    h = rng.standard_normal((b, L, H)).astype(np.float32)
    mask = (rng.random((b, L)) < 0.7).astype(np.int64)
    mask[:, 0] = 1
    mask[1] = 0                                   # all-masked row: mask_sum == 0 -> zeros (empty text, minilm.rs:1123-1125)
    h[2, 3, 10] = np.nan                          # NaN in a live position -> that component scrubbed to 0
    h[3, 5, 20] = np.inf                          # +Inf -> scrubbed
    h[3, 6, 21] = -np.inf
    h[4, 2, 30], h[4, 4, 30] = np.inf, -np.inf    # Inf - Inf = NaN in the running sum -> scrubbed
    h[5] *= 1e-12                                 # norm far below f32::EPSILON: no divide, values returned as pooled
    h[6] *= np.float32(3e-9)                      # norm just around 1.19e-7
    h[7] *= 1e30                                  # squares overflow: norm = +Inf -> divide gives 0 (and Inf/Inf never: values finite)
    mask[8] = 1
    mask[8, 7] = 2                                # only `== 1` counts (minilm.rs:962-966)
    out = sg.pool_finalize(h, mask)
    for i in range(b):
        ref = oracle.pool_finalize(h[i], mask[i])
        assert np.array_equal(out[i].view(np.uint32), ref.view(np.uint32)), i
    assert not out[1].any()
    n5 = float(np.linalg.norm(out[5].astype(np.float64)))
    assert n5 < 1.2e-7                            # the EPSILON branch really was taken
    assert abs(float(np.linalg.norm(out[0].astype(np.float64))) - 1.0) < 1e-5
