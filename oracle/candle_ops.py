"""candle 0.8.x primitive-op semantics restated in numpy f32 (TEST INFRASTRUCTURE).

candle-core / candle-nn are crates.io dependencies of the reference
(/root/reference/Cargo.toml:19-21, "^0.8.2", un-vendored), so each function below
restates the published CPU algorithm of the named candle op; the reference call
site that reaches it is cited.  All arithmetic is f32 (candle CPU F32 path, the
only dtype the reference's CPU build can execute: SURVEY.md section 7 hard part 2).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def linear(x: np.ndarray, w: np.ndarray, b: np.ndarray | None = None) -> np.ndarray:
    """candle_nn::Linear::forward: y = x . W^T (+ b); W is [out, in] row-major.
    Call sites: models/embeddings.rs:131-144 (BERT q/k/v), candle llama/mistral/qwen2 projections."""
    y = np.matmul(x.astype(F32, copy=False), w.astype(F32, copy=False).T)
    if b is not None:
        y = y + b.astype(F32, copy=False)
    return y.astype(F32, copy=False)


def embedding(table: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """candle_nn::Embedding::forward = index_select on dim 0 (models/embeddings.rs:371-372)."""
    return table[ids.astype(np.int64)]


def rms_norm(x: np.ndarray, w: np.ndarray, eps: float) -> np.ndarray:
    """candle_nn::ops::rms_norm (CPU): m = sqrt(sum(x^2)/n + eps); y = x / m * w, f32 throughout."""
    x = x.astype(F32, copy=False)
    n = x.shape[-1]
    s2 = np.sum(x * x, axis=-1, keepdims=True, dtype=F32)
    m = np.sqrt(s2 / F32(n) + F32(eps)).astype(F32)
    return (x / m * w.astype(F32, copy=False)).astype(F32)


def layer_norm(x: np.ndarray, w: np.ndarray, b: np.ndarray, eps: float) -> np.ndarray:
    """candle_nn::ops::layer_norm (CPU fast path used by candle_nn::LayerNorm::forward):
    one-pass mean = sum/n, var = sum2/n - mean^2, y = (x - mean) * (var + eps)^-1/2 * w + b.
    Call sites: models/embeddings.rs:189,241,376."""
    x = x.astype(F32, copy=False)
    n = F32(x.shape[-1])
    s = np.sum(x, axis=-1, keepdims=True, dtype=F32)
    s2 = np.sum(x * x, axis=-1, keepdims=True, dtype=F32)
    mean = s / n
    var = s2 / n - mean * mean
    inv = (F32(1.0) / np.sqrt(var + F32(eps))).astype(F32)
    return ((x - mean) * inv * w.astype(F32, copy=False) + b.astype(F32, copy=False)).astype(F32)


def softmax_last_dim(x: np.ndarray) -> np.ndarray:
    """candle_nn::ops::softmax_last_dim and the generic candle_nn::ops::softmax(dim=-1)
    (models/embeddings.rs:160-162): max-subtract, exp, sum, divide."""
    x = x.astype(F32, copy=False)
    mx = np.max(x, axis=-1, keepdims=True)
    e = np.exp(x - mx).astype(F32)
    return (e / np.sum(e, axis=-1, keepdims=True, dtype=F32)).astype(F32)


def silu(x: np.ndarray) -> np.ndarray:
    """candle Tensor::silu: x / (1 + exp(-x))."""
    x = x.astype(F32, copy=False)
    return (x / (F32(1.0) + np.exp(-x).astype(F32))).astype(F32)


def gelu_tanh(x: np.ndarray) -> np.ndarray:
    """candle Tensor::gelu() = the tanh approximation (models/embeddings.rs:229-231):
    0.5 x (1 + tanh(sqrt(2/pi) x (1 + 0.044715 x^2)))."""
    x = x.astype(F32, copy=False)
    k = F32(np.sqrt(2.0 / np.pi))
    inner = k * x * (F32(1.0) + F32(0.044715) * x * x)
    return (F32(0.5) * x * (F32(1.0) + np.tanh(inner).astype(F32))).astype(F32)


def rope_tables(head_dim: int, max_pos: int, theta: float, theta_pow_f64: bool):
    """cos/sin tables [max_pos, head_dim/2] in f32.
    Llama (candle llama.rs Cache::new): theta is f32 and inv_freq = 1f32 / theta.powf(i as f32 / d as f32).
    Mistral/Qwen2 (RotaryEmbedding::new): theta is f64 and inv_freq = 1f32 / theta.powf(i as f64 / d as f64) as f32.
    Then freqs = positions(f32) outer inv_freq (f32 matmul), cos/sin in f32."""
    i = np.arange(0, head_dim, 2)
    # libm powf / cosf / sinf are (almost always) correctly rounded; model them as f64 evaluation rounded to f32,
    # which is also what the library does on the host (finalize() in fastllm_b200/csrc/fl_lib.cu), so both sides hold identical tables.
    if theta_pow_f64:
        p = np.power(np.float64(theta), i.astype(np.float64) / np.float64(head_dim)).astype(F32)
    else:
        e = (i.astype(F32) / F32(head_dim)).astype(F32)
        p = np.power(np.float64(F32(theta)), e.astype(np.float64)).astype(F32)
    inv = (F32(1.0) / p).astype(F32)
    pos = np.arange(max_pos, dtype=F32)[:, None]
    freqs = (pos * inv[None, :]).astype(F32)
    return np.cos(freqs.astype(np.float64)).astype(F32), np.sin(freqs.astype(np.float64)).astype(F32)


def rope_rotate_half(x: np.ndarray, cos: np.ndarray, sin: np.ndarray) -> np.ndarray:
    """candle_nn::rotary_emb::rope (non-interleaved): x = [x1 | x2] halves of the last dim,
    y1 = x1 cos - x2 sin, y2 = x1 sin + x2 cos.  x: [b, h, t, d]; cos/sin: [t, d/2]."""
    d2 = x.shape[-1] // 2
    x1, x2 = x[..., :d2], x[..., d2:]
    return np.concatenate([x1 * cos - x2 * sin, x1 * sin + x2 * cos], axis=-1).astype(F32)


def repeat_kv(x: np.ndarray, n_rep: int) -> np.ndarray:
    """candle-transformers utils::repeat_kv: [b, nkv, L, d] -> [b, nkv*n_rep, L, d]; q head h uses kv head h // n_rep."""
    if n_rep == 1:
        return x
    return np.repeat(x, n_rep, axis=1)
