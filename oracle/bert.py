"""BERT / MiniLM sentence encoder restated in numpy f32 (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows the reference's hand-written encoder line by line:
  embed_tokens   src/models/embeddings.rs:370-378  word + position embeddings (NO token-type), LayerNorm eps 1e-12 literal (:317)
  BertAttention  src/models/embeddings.rs:130-191  q/k/v linear+bias, scores = q k^T / sqrt(d) (after the matmul), generic softmax,
                                                  NO attention/padding mask, output dense, post-LN residual LN(x + attn)
  BertLayer      src/models/embeddings.rs:222-242  intermediate dense -> .gelu() (candle's TANH approximation) -> output dense,
                                                  post-LN residual LN(x + ffn), eps = config.layer_norm_eps
  mean_pooling   src/models/embeddings.rs:346-368  sum(mask * h) / (mask_count * hidden)   [sic: the divisor is count*hidden]
  normalize_l2   src/models/embeddings.rs:341-344  v / sqrt(sum v^2)  (cancels the pooling divisor)
  embed          src/models/embeddings.rs:397-447  position ids 0..n-1, f32 [hidden] out
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import candle_ops as ops
from . import synth

F32 = np.float32


@dataclass
class BertConfig:
    hidden_size: int = 384
    num_attention_heads: int = 12
    num_hidden_layers: int = 6
    intermediate_size: int = 1536
    max_position_embeddings: int = 512
    layer_norm_eps: float = 1e-12
    vocab_size: int = 30522          # tokenizer.get_vocab_size(false) (embeddings.rs:301-306)


MINILM_L6 = BertConfig()


def tensor_names(cfg: BertConfig):
    H, I = cfg.hidden_size, cfg.intermediate_size
    out = [("embeddings.word_embeddings.weight", (cfg.vocab_size, H), "w"),
           ("embeddings.position_embeddings.weight", (cfg.max_position_embeddings, H), "w"),
           ("embeddings.LayerNorm.weight", (H,), "norm"), ("embeddings.LayerNorm.bias", (H,), "w")]
    for i in range(cfg.num_hidden_layers):
        p = f"encoder.layer.{i}."
        for n, shp in (("attention.self.query", (H, H)), ("attention.self.key", (H, H)), ("attention.self.value", (H, H)),
                       ("attention.output.dense", (H, H)), ("intermediate.dense", (I, H)), ("output.dense", (H, I))):
            out += [(p + n + ".weight", shp, "w"), (p + n + ".bias", (shp[0],), "w")]
        for n in ("attention.output.LayerNorm", "output.LayerNorm"):
            out += [(p + n + ".weight", (H,), "norm"), (p + n + ".bias", (H,), "w")]
    return out


def synth_weights(cfg: BertConfig, seed: int = 0, std: float = 0.02) -> dict:
    w = {}
    for name, shape, kind in tensor_names(cfg):
        w[name] = np.ones(shape, dtype=F32) if kind == "norm" else synth.normal(seed, name, shape, std)
    return w


class MiniLM:
    def __init__(self, cfg: BertConfig, weights: dict):
        self.cfg = cfg
        self.w = {k: np.ascontiguousarray(v, dtype=F32) for k, v in weights.items()}

    def _lin(self, x, name):
        return ops.linear(x, self.w[name + ".weight"], self.w[name + ".bias"])

    def _ln(self, x, name, eps):
        return ops.layer_norm(x, self.w[name + ".weight"], self.w[name + ".bias"], eps)

    def embed_tokens(self, ids: np.ndarray, pos_ids: np.ndarray) -> np.ndarray:
        e = ops.embedding(self.w["embeddings.word_embeddings.weight"], ids) + \
            ops.embedding(self.w["embeddings.position_embeddings.weight"], pos_ids)
        return self._ln(e.astype(F32), "embeddings.LayerNorm", 1e-12)

    def attention(self, li: int, h: np.ndarray) -> np.ndarray:
        cfg = self.cfg
        b, t, H = h.shape
        nh = cfg.num_attention_heads
        d = H // nh
        p = f"encoder.layer.{li}.attention."
        split = lambda x: x.reshape(b, t, nh, d).transpose(0, 2, 1, 3)
        q, k, v = split(self._lin(h, p + "self.query")), split(self._lin(h, p + "self.key")), split(self._lin(h, p + "self.value"))
        scores = np.matmul(q, k.transpose(0, 1, 3, 2)).astype(F32)
        scores = (scores / F32(np.sqrt(np.float64(d)))).astype(F32)
        probs = ops.softmax_last_dim(scores)
        ctx = np.matmul(probs, v).astype(F32).transpose(0, 2, 1, 3).reshape(b, t, H)
        out = self._lin(ctx, p + "output.dense")
        return self._ln((h + out).astype(F32), p + "output.LayerNorm", cfg.layer_norm_eps)

    def layer(self, li: int, h: np.ndarray) -> np.ndarray:
        p = f"encoder.layer.{li}."
        h = self.attention(li, h)
        inter = ops.gelu_tanh(self._lin(h, p + "intermediate.dense"))
        out = self._lin(inter, p + "output.dense")
        return self._ln((h + out).astype(F32), p + "output.LayerNorm", self.cfg.layer_norm_eps)

    def forward(self, ids: np.ndarray, pos_ids: np.ndarray | None = None) -> np.ndarray:
        ids = np.asarray(ids)
        if pos_ids is None:
            pos_ids = np.broadcast_to(np.arange(ids.shape[1]), ids.shape)
        h = self.embed_tokens(ids, pos_ids)
        for li in range(self.cfg.num_hidden_layers):
            h = self.layer(li, h)
        return h

    @staticmethod
    def mean_pooling(h: np.ndarray, mask: np.ndarray) -> np.ndarray:
        hidden = h.shape[-1]
        m = np.broadcast_to(mask[:, :, None].astype(F32), h.shape)
        summed = np.sum(h * m, axis=1, dtype=F32)
        n_tokens = np.sum(np.sum(m, axis=1, dtype=F32), axis=1, dtype=F32)[:, None]   # = count * hidden  [sic]
        assert hidden > 0
        return (summed / n_tokens).astype(F32)

    @staticmethod
    def normalize_l2(v: np.ndarray) -> np.ndarray:
        return (v / np.sqrt(np.sum(v * v, axis=1, keepdims=True, dtype=F32))).astype(F32)

    def embed_ids(self, ids: np.ndarray, mask: np.ndarray | None = None) -> np.ndarray:
        """ids u32 [b, t] (+ mask [b, t], default all ones as the tokenizer returns for a single sentence) -> f32 [b, H]."""
        ids = np.asarray(ids)
        if mask is None:
            mask = np.ones(ids.shape, dtype=np.uint32)
        return self.normalize_l2(self.mean_pooling(self.forward(ids), np.asarray(mask)))
