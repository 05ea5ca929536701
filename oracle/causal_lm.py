"""Llama / Mistral / Qwen2 causal-LM forward + the reference adapters' offset rules + the
greedy generate loop, restated in numpy f32 (TEST INFRASTRUCTURE, see oracle/__init__.py).

What follows which reference line:
  * LlamaAdapter.forward      -> src/models/llama.rs:147-149  (pos comes from the caller; KV lives in the cache)
  * MistralAdapter.forward    -> src/models/mistral.rs:206-236 (ignores pos; clears KV when offset==0; offset += 1 PER CALL)
  * QwenAdapter.forward       -> src/models/qwen.rs:129-145    (same rule as Mistral)
  * generate()                -> src/models/mod.rs:363-463     (prefill at pos 0, arg-max, EOS break before emit)
  * sample_argmax()           -> candle_transformers::generation::LogitsProcessor (ties -> LAST index)
  * CausalLM.forward          -> candle-transformers 0.8.x models::{llama,mistral,qwen2} (un-vendored; SURVEY.md section 8a rows 1,3,4)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import candle_ops as ops
from . import synth

F32 = np.float32

ARCH_LLAMA, ARCH_MISTRAL, ARCH_QWEN2, ARCH_MIXTRAL = "llama", "mistral", "qwen2", "mixtral"


@dataclass
class CausalLMConfig:
    arch: str
    hidden_size: int
    intermediate_size: int
    vocab_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    max_position_embeddings: int = 4096
    sliding_window: int = 4096          # Mistral/Qwen2 adapters pass unwrap_or(4096) (mistral.rs:139, qwen.rs:49)
    qkv_bias: bool = False              # Qwen2 only
    num_local_experts: int = 0          # Mixtral only
    num_experts_per_tok: int = 2

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    def validate(self):
        # models/config.rs:20-54, mistral.rs:109-127: the reference panics on these
        assert self.head_dim * self.num_attention_heads == self.hidden_size, "hidden_size must be divisible by num_attention_heads"
        assert self.head_dim % 2 == 0, "head_dim must be even for RoPE embeddings"
        assert self.num_attention_heads % self.num_key_value_heads == 0, "num_attention_heads must be divisible by num_key_value_heads"


TINYLLAMA = CausalLMConfig(ARCH_LLAMA, 2048, 5632, 32000, 22, 32, 4, 1e-5, 1e4, 2048)
MISTRAL_7B = CausalLMConfig(ARCH_MISTRAL, 4096, 14336, 32000, 32, 32, 8, 1e-5, 1e4, 32768, 4096)
QWEN25_7B = CausalLMConfig(ARCH_QWEN2, 3584, 18944, 152064, 28, 28, 4, 1e-6, 1e6, 32768, 4096, qkv_bias=True)
MIXTRAL_8X7B = CausalLMConfig(ARCH_MIXTRAL, 4096, 14336, 32000, 32, 32, 8, 1e-5, 1e6, 32768, 4096,
                              num_local_experts=8, num_experts_per_tok=2)


def tensor_names(cfg: CausalLMConfig):
    """(name, shape, kind) for every tensor candle's VarBuilder would request (HF naming)."""
    H, I, V, d = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.head_dim
    nq, nkv = cfg.num_attention_heads * d, cfg.num_key_value_heads * d
    out = [("model.embed_tokens.weight", (V, H), "w")]
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        out += [(p + "input_layernorm.weight", (H,), "norm"),
                (p + "self_attn.q_proj.weight", (nq, H), "w"),
                (p + "self_attn.k_proj.weight", (nkv, H), "w"),
                (p + "self_attn.v_proj.weight", (nkv, H), "w")]
        if cfg.qkv_bias:
            out += [(p + "self_attn.q_proj.bias", (nq,), "w"),
                    (p + "self_attn.k_proj.bias", (nkv,), "w"),
                    (p + "self_attn.v_proj.bias", (nkv,), "w")]
        out += [(p + "self_attn.o_proj.weight", (H, nq), "w"),
                (p + "post_attention_layernorm.weight", (H,), "norm")]
        if cfg.arch == ARCH_MIXTRAL:
            out += [(p + "block_sparse_moe.gate.weight", (cfg.num_local_experts, H), "w")]
            for e in range(cfg.num_local_experts):
                q = p + f"block_sparse_moe.experts.{e}."
                out += [(q + "w1.weight", (I, H), "w"), (q + "w2.weight", (H, I), "w"), (q + "w3.weight", (I, H), "w")]
        else:
            out += [(p + "mlp.gate_proj.weight", (I, H), "w"),
                    (p + "mlp.up_proj.weight", (I, H), "w"),
                    (p + "mlp.down_proj.weight", (H, I), "w")]
    out += [("model.norm.weight", (H,), "norm"), ("lm_head.weight", (V, H), "w")]
    return out


def synth_weights(cfg: CausalLMConfig, seed: int = 0, std: float = 0.02) -> dict:
    """Random-init weights (bf16-rounded values held as f32): weights/biases ~N(0, std^2), norm weights 1.0."""
    w = {}
    for name, shape, kind in tensor_names(cfg):
        w[name] = np.ones(shape, dtype=F32) if kind == "norm" else synth.normal(seed, name, shape, std)
    return w


class CausalLM:
    """One candle model instance: weights + (for Mistral/Qwen2) the model-internal KV cache."""

    def __init__(self, cfg: CausalLMConfig, weights: dict, rope_len: int | None = None, kv_dtype: str = "f32"):
        """kv_dtype="bf16" rounds K (after RoPE) and V to bf16 before they enter the cache -- what the reference's
        server build does implicitly (KV in model dtype BF16, main.rs:120) and what the product's paged cache stores.
        The default "f32" is the pure candle-CPU-F32 restatement."""
        cfg.validate()
        self.cfg = cfg
        self.kv_dtype = kv_dtype
        self.w = {k: np.ascontiguousarray(v, dtype=F32) for k, v in weights.items()}
        n = rope_len or cfg.max_position_embeddings
        # Llama: f32 theta pow; Mistral/Qwen2/Mixtral: f64 theta pow then f32 (see candle_ops.rope_tables)
        self.cos, self.sin = ops.rope_tables(cfg.head_dim, n, cfg.rope_theta, cfg.arch != ARCH_LLAMA)
        self.clear_kv_cache()

    def clear_kv_cache(self):
        self.kv = [None] * self.cfg.num_hidden_layers

    @property
    def kv_len(self) -> int:
        return 0 if self.kv[0] is None else self.kv[0][0].shape[2]

    # -- attention -------------------------------------------------------------------------------
    def _mask(self, t: int, offset: int) -> np.ndarray | None:
        """Additive mask [t, offset+t] (Mistral/Qwen2 prepare_decoder_attention_mask) or the Llama
        t x t masked_fill pattern; only built when t > 1."""
        if t <= 1:
            return None
        i = np.arange(t)[:, None]
        j = np.arange(t)[None, :]
        if self.cfg.arch == ARCH_LLAMA:
            # candle llama.rs Cache::mask: mask[i][j] = (j > i); masked_fill(att, mask, -inf).  t x t only,
            # so a Llama prefill must start from an empty cache.
            assert offset == 0, "candle Llama cannot prefill t>1 on a non-empty cache (t x t mask)"
            banned = j > i
        else:
            banned = (i < j) | (j + self.cfg.sliding_window < i)
        m = np.where(banned, F32(-np.inf), F32(0.0)).astype(F32)
        if offset > 0:
            m = np.concatenate([np.zeros((t, offset), dtype=F32), m], axis=1)
        return m

    def _attention(self, li: int, x: np.ndarray, rope_offset: int) -> np.ndarray:
        cfg, w = self.cfg, self.w
        b, t, _ = x.shape
        nh, nkv, d = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        p = f"model.layers.{li}.self_attn."
        bias = (lambda n: w[p + n + ".bias"]) if cfg.qkv_bias else (lambda n: None)
        q = ops.linear(x, w[p + "q_proj.weight"], bias("q_proj")).reshape(b, t, nh, d).transpose(0, 2, 1, 3)
        k = ops.linear(x, w[p + "k_proj.weight"], bias("k_proj")).reshape(b, t, nkv, d).transpose(0, 2, 1, 3)
        v = ops.linear(x, w[p + "v_proj.weight"], bias("v_proj")).reshape(b, t, nkv, d).transpose(0, 2, 1, 3)
        cos, sin = self.cos[rope_offset:rope_offset + t], self.sin[rope_offset:rope_offset + t]
        q = ops.rope_rotate_half(q, cos, sin)
        k = ops.rope_rotate_half(k, cos, sin)
        if self.kv_dtype == "bf16":
            k, v = synth.round_bf16(k), synth.round_bf16(v)
        prev = self.kv[li]
        kv_before = 0 if prev is None else prev[0].shape[2]
        if prev is not None:                       # Tensor::cat(&[prev, new], 2)  (K6)
            k = np.concatenate([prev[0], k], axis=2)
            v = np.concatenate([prev[1], v], axis=2)
        self.kv[li] = (k, v)
        kk, vv = ops.repeat_kv(k, nh // nkv), ops.repeat_kv(v, nh // nkv)
        att = np.matmul(q, kk.transpose(0, 1, 3, 2)).astype(F32)
        if cfg.arch == ARCH_LLAMA:
            att = (att / F32(np.sqrt(np.float64(d)))).astype(F32)          # att / sqrt(d)
        else:
            att = (att * F32(1.0 / np.sqrt(np.float64(d)))).astype(F32)    # att * (1/sqrt(d)) via affine
        m = self._mask(t, kv_before)
        if m is not None:
            att = att + m if cfg.arch != ARCH_LLAMA else np.where(np.isneginf(m), F32(-np.inf), att)
        att = ops.softmax_last_dim(att)
        y = np.matmul(att, vv).astype(F32).transpose(0, 2, 1, 3).reshape(b, t, nh * d)
        return ops.linear(y, w[p + "o_proj.weight"])

    # -- MLP / MoE -------------------------------------------------------------------------------
    def _mlp(self, li: int, x: np.ndarray) -> np.ndarray:
        w, p = self.w, f"model.layers.{li}.mlp."
        return ops.linear(ops.silu(ops.linear(x, w[p + "gate_proj.weight"])) * ops.linear(x, w[p + "up_proj.weight"]),
                          w[p + "down_proj.weight"])

    def _moe(self, li: int, x: np.ndarray) -> np.ndarray:
        from .mixtral import sparse_moe_block
        p = f"model.layers.{li}.block_sparse_moe."
        E = self.cfg.num_local_experts
        experts = [(self.w[p + f"experts.{e}.w1.weight"], self.w[p + f"experts.{e}.w2.weight"],
                    self.w[p + f"experts.{e}.w3.weight"]) for e in range(E)]
        return sparse_moe_block(x, self.w[p + "gate.weight"], experts, self.cfg.num_experts_per_tok)

    # -- full forward ----------------------------------------------------------------------------
    def forward(self, ids: np.ndarray, rope_offset: int) -> np.ndarray:
        """ids u32 [b, t] -> f32 logits [b, V] of the LAST position (K17)."""
        cfg, w = self.cfg, self.w
        ids = np.asarray(ids)
        assert ids.ndim == 2
        x = ops.embedding(w["model.embed_tokens.weight"], ids)
        for li in range(cfg.num_hidden_layers):
            p = f"model.layers.{li}."
            h = ops.rms_norm(x, w[p + "input_layernorm.weight"], cfg.rms_norm_eps)
            x = (self._attention(li, h, rope_offset) + x).astype(F32)
            h = ops.rms_norm(x, w[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
            ff = self._moe(li, h) if cfg.arch == ARCH_MIXTRAL else self._mlp(li, h)
            x = (ff + x).astype(F32)
        last = x[:, -1, :]
        last = ops.rms_norm(last, w["model.norm.weight"], cfg.rms_norm_eps)
        return ops.linear(last, w["lm_head.weight"])


# ---- reference adapters (the L2 layer of SURVEY.md section 1) -------------------------------------

@dataclass
class OffsetCache:
    """ModelCache (src/models/cache.rs:5-46): only a sequence offset."""
    seqlen_offset: int = 0

    def increment_offset(self):
        self.seqlen_offset += 1

    def reset(self):
        self.seqlen_offset = 0

    def get_offset(self) -> int:
        return self.seqlen_offset


class LlamaAdapter:
    """LlamaWithConfig (llama.rs:147-149): forward(input, pos, cache) -> model.forward(input, pos, &mut cache.inner).
    The KV lives in the cache; a fresh cache is built per generate (mod.rs:370)."""

    def __init__(self, model: CausalLM):
        self.model = model

    def initialize_cache(self) -> OffsetCache:
        self.model.clear_kv_cache()
        return OffsetCache()

    def forward(self, ids, pos: int, cache: OffsetCache) -> np.ndarray:
        return self.model.forward(ids, pos)                     # [b, V] f32


class MistralAdapter:
    """MistralWithConfig / QwenWithConfig (mistral.rs:206-236, qwen.rs:129-145): `_pos` ignored; KV cleared when
    cache.seqlen_offset == 0; RoPE offset = cache.seqlen_offset, which then grows by ONE per call.  `faithful=False`
    gives the position-correct variant (offset += t) for the HF cross-check."""

    def __init__(self, model: CausalLM, faithful: bool = True):
        self.model, self.faithful = model, faithful

    def initialize_cache(self) -> OffsetCache:
        return OffsetCache()

    def forward(self, ids, _pos: int, cache: OffsetCache) -> np.ndarray:
        if cache.get_offset() == 0:
            self.model.clear_kv_cache()
        out = self.model.forward(ids, cache.get_offset())
        if self.faithful:
            cache.increment_offset()
        else:
            cache.seqlen_offset += np.asarray(ids).shape[1]
        return out[:, None, :]                                   # [b, 1, V]


QwenAdapter = MistralAdapter


def make_adapter(model: CausalLM, faithful: bool = True):
    return LlamaAdapter(model) if model.cfg.arch == ARCH_LLAMA else MistralAdapter(model, faithful)


def sample_argmax(logits: np.ndarray) -> int:
    """LogitsProcessor::sample_argmax: iter().enumerate().max_by(total_cmp) => the LAST index among equal maxima."""
    v = np.asarray(logits, dtype=F32).reshape(-1)
    m = v.max()
    return int(np.flatnonzero(v == m)[-1])


def generate(adapter, prompt_ids, max_tokens: int, eos_id: int | None = 2, return_logits: bool = False, temperature: float = 0.0):
    """Model::generate (mod.rs:363-463): fresh cache, LogitsProcessor::new(0, Some(temperature as f64), None) (arg-max below
    1e-7, the API default), prefill [1, N] at pos 0, then per step: logits.get(0).flatten_all -> sample -> break on EOS
    *before* emitting -> forward([1,1], pos); pos += 1."""
    from .sampling import LogitsProcessor
    logits_processor = LogitsProcessor(0, float(np.float32(temperature)))
    cache = adapter.initialize_cache()
    ids = np.asarray(prompt_ids, dtype=np.uint32).reshape(1, -1)
    pos = 0
    logits = adapter.forward(ids, pos, cache)
    pos += ids.shape[1]
    out, all_logits = [], []
    for _ in range(max_tokens):
        last = np.asarray(logits)[0].reshape(-1)
        all_logits.append(last.copy())
        tok = logits_processor.sample(last)
        if eos_id is not None and tok == eos_id:
            break
        out.append(tok)
        logits = adapter.forward(np.array([[tok]], dtype=np.uint32), pos, cache)
        pos += 1
    return (out, all_logits) if return_logits else out
