"""Generate tests/golden/*.npz and cross-check the oracle against HuggingFace transformers (torch CPU f32).

Run here (authoring container, no GPU):   python tests/golden/make_golden.py

The reference has no golden vectors for the forward pass and cannot be built (no Rust
toolchain), so the pins are: (1) the oracle, which restates candle 0.8.x; (2) an independent implementation, HF
transformers, with the known deltas neutralised (BERT: token-type embedding zeroed, tanh-GELU, all-ones mask;
Mistral/Qwen2: compared in POSITION-CORRECT mode because HF does not have the adapters' +1-per-call offset quirk).
This script asserts (1) == (2) to 2e-4 max-abs on logits / embeddings and identical greedy ids, then freezes the
oracle's outputs (both reference-faithful and position-correct modes) as fixtures.  Weights are not stored: they
regenerate bit-exactly from (seed, std) via oracle/synth.py.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import bert as obert  # noqa: E402
from oracle import causal_lm as ocl  # noqa: E402
from oracle import mixtral as omix  # noqa: E402
from oracle import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
STD = 0.08
TOL = 2e-4

TINY = {
    "llama": ocl.CausalLMConfig("llama", 64, 176, 256, 2, 4, 2, 1e-5, 1e4, 128),
    "llama_gqa8": ocl.CausalLMConfig("llama", 256, 352, 512, 2, 8, 1, 1e-5, 1e4, 128),
    "mistral": ocl.CausalLMConfig("mistral", 128, 224, 320, 2, 4, 2, 1e-5, 1e4, 256, 4096),
    "mistral_sw": ocl.CausalLMConfig("mistral", 64, 96, 128, 2, 4, 2, 1e-5, 1e4, 256, 5),
    "qwen2": ocl.CausalLMConfig("qwen2", 112, 160, 288, 2, 7, 1, 1e-6, 1e6, 256, 4096, qkv_bias=True),
    "mixtral": ocl.CausalLMConfig("mixtral", 64, 96, 128, 2, 4, 2, 1e-5, 1e6, 256, 4096, num_local_experts=4,
                                  num_experts_per_tok=2),
}
PROMPT, NEW = 12, 8


def hf_model(cfg: ocl.CausalLMConfig, w: dict):
    import torch
    import transformers as tf
    common = dict(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size, vocab_size=cfg.vocab_size,
                  num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                  num_key_value_heads=cfg.num_key_value_heads, rms_norm_eps=cfg.rms_norm_eps,
                  max_position_embeddings=cfg.max_position_embeddings, tie_word_embeddings=False,
                  attn_implementation="eager")
    rope = dict(rope_parameters={"rope_type": "default", "rope_theta": float(cfg.rope_theta)})
    sd = {k: torch.from_numpy(v.copy()) for k, v in w.items()}
    if cfg.arch == "llama":
        m = tf.LlamaForCausalLM(tf.LlamaConfig(**common, **rope))
    elif cfg.arch == "mistral":
        # candle bans j + sw < i (window of sw+1 keys); HF keeps i - j < sliding_window  =>  HF window = sw + 1
        m = tf.MistralForCausalLM(tf.MistralConfig(**common, **rope, sliding_window=cfg.sliding_window + 1))
    elif cfg.arch == "qwen2":
        m = tf.Qwen2ForCausalLM(tf.Qwen2Config(**common, **rope, use_sliding_window=False))
    else:
        m = tf.MixtralForCausalLM(tf.MixtralConfig(**common, **rope, num_local_experts=cfg.num_local_experts,
                                                   num_experts_per_tok=cfg.num_experts_per_tok, sliding_window=None))
        E = cfg.num_local_experts
        for li in range(cfg.num_hidden_layers):
            p = f"model.layers.{li}."
            sd[p + "mlp.gate.weight"] = sd.pop(p + "block_sparse_moe.gate.weight")
            w1 = [sd.pop(p + f"block_sparse_moe.experts.{e}.w1.weight") for e in range(E)]
            w2 = [sd.pop(p + f"block_sparse_moe.experts.{e}.w2.weight") for e in range(E)]
            w3 = [sd.pop(p + f"block_sparse_moe.experts.{e}.w3.weight") for e in range(E)]
            sd[p + "mlp.experts.gate_up_proj"] = torch.stack([torch.cat([a, b], 0) for a, b in zip(w1, w3)])
            sd[p + "mlp.experts.down_proj"] = torch.stack(w2)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in k or "inv_freq" in k for k in missing), missing
    return m.eval().float()


def hf_generate(m, prompt: np.ndarray, n_new: int):
    import torch
    with torch.no_grad():
        ids = torch.from_numpy(prompt.astype(np.int64))[None]
        out = m(ids, use_cache=True)
        past, logits = out.past_key_values, [out.logits[0, -1].numpy().copy()]
        toks = []
        for _ in range(n_new):
            t = ocl.sample_argmax(logits[-1])
            toks.append(t)
            out = m(torch.tensor([[t]]), past_key_values=past, use_cache=True)
            past = out.past_key_values
            logits.append(out.logits[0, -1].numpy().copy())
    return toks, logits[:-1]


def make_causal(name: str, cfg: ocl.CausalLMConfig, check_hf: bool = True):
    seed = 7
    w = ocl.synth_weights(cfg, seed, STD)
    rng = np.random.default_rng(11)
    for k in w:                       # non-trivial norm weights so the norm multiply is exercised
        if k.endswith("norm.weight"):
            w[k] = synth.round_bf16((1.0 + 0.1 * rng.standard_normal(w[k].shape)).astype(np.float32))
    prompt = synth.token_ids(1, cfg.vocab_size, (PROMPT,))
    model = ocl.CausalLM(cfg, w)
    res = {"prompt": prompt, "seed": seed, "std": STD}
    for mode, faithful in (("faithful", True), ("poscorrect", False)):
        if cfg.arch == "llama" and not faithful:
            continue
        toks, logits = ocl.generate(ocl.make_adapter(model, faithful), prompt, NEW, eos_id=None, return_logits=True)
        res[f"{mode}_ids"] = np.array(toks, dtype=np.uint32)
        res[f"{mode}_logits"] = np.stack(logits).astype(np.float32)
    for k in w:
        if k.endswith("norm.weight"):
            res["w:" + k] = w[k]
    if check_hf:
        hm = hf_model(cfg, w)
        htoks, hlogits = hf_generate(hm, prompt, NEW)
        key = "faithful" if cfg.arch == "llama" else "poscorrect"
        if cfg.sliding_window < PROMPT + NEW:
            # candle applies the window only inside the PREFILL mask and never trims the KV cache, so decode attends the
            # full history (SURVEY.md section 5 "Long context"); HF windows decode too.  Only the prefill logits compare.
            htoks, hlogits = list(res[f"{key}_ids"]), hlogits[:1]
        err = max(float(np.abs(a - b).max()) for a, b in zip(hlogits, res[f"{key}_logits"]))
        assert htoks == list(res[f"{key}_ids"]), (name, htoks, res[f"{key}_ids"])
        assert err < TOL, (name, err)
        print(f"{name}: oracle == HF transformers, max-abs logits err {err:.2e}, ids {htoks}")
    np.savez(os.path.join(OUT, f"causal_{name}.npz"), **res)


def make_bert():
    import torch
    import transformers as tf
    cfg = obert.BertConfig(128, 4, 2, 256, 64, 1e-12, 200)     # head_dim 32 like all-MiniLM-L6-v2
    seed = 5
    w = obert.synth_weights(cfg, seed, STD)
    rng = np.random.default_rng(3)
    for k in w:
        if k.endswith("LayerNorm.weight"):
            w[k] = synth.round_bf16((1.0 + 0.1 * rng.standard_normal(w[k].shape)).astype(np.float32))
    ids = synth.token_ids(2, cfg.vocab_size, (3, 16))
    emb = obert.MiniLM(cfg, w).embed_ids(ids)
    hidden = obert.MiniLM(cfg, w).forward(ids)
    hc = tf.BertConfig(hidden_size=128, num_attention_heads=4, num_hidden_layers=2, intermediate_size=256,
                       max_position_embeddings=64, layer_norm_eps=1e-12, vocab_size=200,
                       hidden_act="gelu_pytorch_tanh", attn_implementation="eager")
    hm = tf.BertModel(hc, add_pooling_layer=False).eval().float()
    sd = {k: torch.from_numpy(v.copy()) for k, v in w.items()}
    sd["embeddings.token_type_embeddings.weight"] = torch.zeros(2, 128)
    missing, unexpected = hm.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k or "token_type_ids" in k for k in missing), (missing, unexpected)
    with torch.no_grad():
        hh = hm(torch.from_numpy(ids.astype(np.int64))).last_hidden_state.numpy()
    err = float(np.abs(hh - hidden).max())
    assert err < TOL, err
    print(f"bert: oracle == HF transformers, max-abs hidden err {err:.2e}")
    res = {"ids": ids, "seed": seed, "std": STD, "embeddings": emb, "hidden": hidden}
    for k in w:
        if k.endswith("LayerNorm.weight"):
            res["w:" + k] = w[k]
    np.savez(os.path.join(OUT, "bert_tiny.npz"), **res)


def make_router():
    """Mixtral router known-answer vectors including exact ties (lower expert index must win)."""
    logits = np.array([[0.1, 0.7, 0.7, -1.0, 0.0, 0.2, 0.7, 0.1],
                       [1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0],
                       [-3.0, 2.0, 0.5, 2.5, -0.5, 2.5, 0.0, 1.0],
                       [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 5.0]], dtype=np.float32)
    idx, wts = omix.route_top_k(logits, 2)
    assert idx.tolist() == [[1, 2], [0, 1], [3, 5], [7, 0]], idx
    np.savez(os.path.join(OUT, "mixtral_router.npz"), logits=logits, idx=idx, wts=wts)
    print("router KAT:", idx.tolist())


if __name__ == "__main__":
    for n, c in TINY.items():
        make_causal(n, c)
    make_bert()
    make_router()
