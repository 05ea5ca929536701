"""GPU parity of the BERT/MiniLM encoder path (tcgen05 GEMMs + fused attention) against the oracle.

Tolerance (north_star): cosine >= 0.999 per embedding; the path runs bf16 tensor-core inputs with f32 accumulation, so
max-abs on the unit-norm embedding components is also held to 2e-2.
"""
import os

import numpy as np
import pytest

from oracle import bert as obert
from oracle import synth

pytestmark = pytest.mark.gpu
COS_TOL = 0.999


def _cos(a, b):
    return (a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))


def _golden(golden_dir):
    g = np.load(f"{golden_dir}/bert_tiny.npz")
    cfg = obert.BertConfig(128, 4, 2, 256, 64, 1e-12, 200)
    w = obert.synth_weights(cfg, int(g["seed"]), float(g["std"]))
    for k in g.files:
        if k.startswith("w:"):
            w[k[2:]] = g[k]
    return g, cfg, w


def _product(cfg, w, **kw):
    from fastllm_b200 import models
    pc = models.BertConfig(cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers, cfg.intermediate_size,
                           cfg.max_position_embeddings, cfg.layer_norm_eps, cfg.vocab_size)
    return models.MiniLMModel(pc, w, 0, **kw)


def test_bert_golden_embeddings(golden_dir):
    g, cfg, w = _golden(golden_dir)
    m = _product(cfg, w)
    emb = m.embed_ids(g["ids"])
    cos = _cos(emb, g["embeddings"])
    print("tiny bert: min cosine", cos.min(), "max-abs", np.abs(emb - g["embeddings"]).max())
    assert cos.min() >= COS_TOL
    assert np.abs(emb - g["embeddings"]).max() <= 2e-2
    assert np.allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-4)
    assert m.embedding_size() == 128 and m.get_family() == "bert" and m.supports_architecture("BertModel")


def test_bert_ragged_lengths_and_mask(golden_dir):
    """Sentence lengths that are not multiples of 16, batch sizes that do not fill a 128-row GEMM tile, and a padding mask
    in the pooling (the reference applies the mask ONLY in mean_pooling, never inside attention)."""
    g, cfg, w = _golden(golden_dir)
    m = _product(cfg, w)
    o = obert.MiniLM(cfg, w)
    for b, t in [(1, 1), (1, 7), (5, 33), (3, 64)]:
        ids = synth.token_ids(b * 100 + t, cfg.vocab_size, (b, t))
        mask = np.ones((b, t), dtype=np.uint32)
        if t > 3:
            mask[:, -2:] = 0
        want = o.embed_ids(ids, mask)
        got = m.embed_ids(ids, mask)
        assert _cos(got, want).min() >= COS_TOL, (b, t)


@pytest.mark.parametrize("b,t", [(1, 129), (3, 200), (2, 256), (1, 384), (2, 512)])
def test_minilm_long_sentences(b, t):
    """The reference enforces no sentence limit below the 512-row position table (embeddings.rs:285-286, 416): sentences of more
    than 128 tokens run the key-tiled attention kernel (online softmax) and must match the oracle like the short ones."""
    cfg = obert.MINILM_L6
    w = obert.synth_weights(cfg, 0, 0.02)
    ids = synth.token_ids(300 + t, cfg.vocab_size, (b, t))
    mask = np.ones((b, t), dtype=np.uint32)
    mask[:, -3:] = 0
    want = obert.MiniLM(cfg, w).embed_ids(ids, mask)
    got = _product(cfg, w).embed_ids(ids, mask)
    cos = _cos(got, want)
    print(f"MiniLM-L6 {b}x{t}: min cosine {cos.min():.6f}, max-abs {np.abs(got - want).max():.3e}")
    assert cos.min() >= COS_TOL and np.abs(got - want).max() <= 2e-2


def test_minilm_true_shape_batch():
    """all-MiniLM-L6-v2 shapes (H=384, 12 heads, I=1536, 6 layers), 16 x 128 tokens, synthetic weights."""
    cfg = obert.MINILM_L6
    w = obert.synth_weights(cfg, 0, 0.02)
    ids = synth.token_ids(2, cfg.vocab_size, (16, 128))
    want = obert.MiniLM(cfg, w).embed_ids(ids)
    m = _product(cfg, w)
    got = m.embed_ids(ids)
    cos = _cos(got, want)
    print("MiniLM-L6 16x128: min cosine", cos.min(), "max-abs", np.abs(got - want).max())
    assert cos.min() >= COS_TOL
    # device-side synthetic init must give the same embeddings as uploading the host-generated weights
    m2 = _product(cfg, None, random_seed=0, std=0.02)
    assert np.array_equal(m2.embed_ids(ids), got)
    # similarity API: cosine of a sentence with itself is 1
    assert abs(m.compute_similarity(ids[0], ids[0]) - 1.0) < 1e-5


def test_bert_errors():
    from fastllm_b200 import FastllmError, models
    cfg = obert.BertConfig(128, 4, 1, 256, 64, 1e-12, 200)
    m = _product(cfg, None, random_seed=1)
    with pytest.raises(FastllmError):
        m.embed_ids(np.full((1, 4), 200, dtype=np.uint32))            # id out of range
    with pytest.raises(FastllmError):
        m.embed_ids(np.ones((1, 65), dtype=np.uint32))                # > max_position_embeddings (64 here): the position lookup fails
    with pytest.raises(FastllmError):
        models.MiniLMModel(models.BertConfig(96, 4, 1, 256, 64, 1e-12, 200), None)   # head_dim 24 unsupported


# The reference's one numeric test (models/embeddings.rs:473-511).  It needs the real all-MiniLM-L6-v2 checkpoint, which the reference
# downloads and this image cannot: opt-in through FASTLLM_MINILM_DIR = a directory holding model.safetensors, tokenizer.json
# (+ config.json) of sentence-transformers/all-MiniLM-L6-v2.  Sentences and bands are the reference's, verbatim.
SIMILARITY_BANDS = [
    ("I really enjoyed the movie. It was a great film.", "The movie was excellent and I had a good time watching it.", 0.8, None),
    ("I enjoy programming in Python because it's easy to read.", "Java is a popular programming language for enterprise applications.",
     0.4, 0.8),
    ("The recipe calls for two cups of flour and one cup of sugar.", "The Hubble telescope has captured stunning images of distant galaxies.",
     None, 0.4),
]


@pytest.mark.skipif(not os.environ.get("FASTLLM_MINILM_DIR"), reason="opt-in: set FASTLLM_MINILM_DIR to a local all-MiniLM-L6-v2 checkout")
def test_embedding_similarities():
    from tokenizers import Tokenizer
    from fastllm_b200 import models, safetensors_io
    d = os.environ["FASTLLM_MINILM_DIR"]
    tok = Tokenizer.from_file(os.path.join(d, "tokenizer.json"))
    tensors = {n.removeprefix("bert."): a for n, a in safetensors_io.iter_tensors(d)}
    m = models.MiniLMModel(models.BertConfig(vocab_size=tok.get_vocab_size(False)), tensors)      # embeddings.rs:301-306
    ids = lambda s: np.asarray(tok.encode(s.lower(), add_special_tokens=True).ids, dtype=np.uint32)   # do_lower_case (:398-403)
    for a, b, lo, hi in SIMILARITY_BANDS:
        sim = m.compute_similarity(ids(a), ids(b))
        assert (lo is None or sim > lo) and (hi is None or sim < hi), (a, b, sim)
