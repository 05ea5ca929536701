"""Shared helpers for the parity tests: tiny golden configs, oracle<->product glue."""
from __future__ import annotations

import os

import numpy as np

from oracle import causal_lm as ocl
from oracle import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TINY = {
    "llama": ocl.CausalLMConfig("llama", 64, 176, 256, 2, 4, 2, 1e-5, 1e4, 128),
    "llama_gqa8": ocl.CausalLMConfig("llama", 256, 352, 512, 2, 8, 1, 1e-5, 1e4, 128),
    "mistral": ocl.CausalLMConfig("mistral", 128, 224, 320, 2, 4, 2, 1e-5, 1e4, 256, 4096),
    "mistral_sw": ocl.CausalLMConfig("mistral", 64, 96, 128, 2, 4, 2, 1e-5, 1e4, 256, 5),
    "qwen2": ocl.CausalLMConfig("qwen2", 112, 160, 288, 2, 7, 1, 1e-6, 1e6, 256, 4096, qkv_bias=True),
    "mixtral": ocl.CausalLMConfig("mixtral", 64, 96, 128, 2, 4, 2, 1e-5, 1e6, 256, 4096, num_local_experts=4,
                                  num_experts_per_tok=2),
}


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN, f"causal_{name}.npz"))


def golden_weights(name: str):
    """Regenerate the fixture's weights bit-exactly: synth(seed, std) + the stored norm weights."""
    g = load_golden(name)
    cfg = TINY[name]
    w = ocl.synth_weights(cfg, int(g["seed"]), float(g["std"]))
    for k in g.files:
        if k.startswith("w:"):
            w[k[2:]] = g[k]
    return cfg, w, g


def product_model(cfg: ocl.CausalLMConfig, weights):
    """Build the product-side adapter (fastllm_b200.models) for an oracle config."""
    from fastllm_b200 import models
    cf = models.ConfigFile(cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.num_hidden_layers,
                           cfg.num_attention_heads, cfg.num_key_value_heads, cfg.rms_norm_eps, cfg.rope_theta,
                           cfg.max_position_embeddings, cfg.sliding_window if cfg.arch != "llama" else None,
                           num_local_experts=cfg.num_local_experts, num_experts_per_tok=cfg.num_experts_per_tok)
    cls = {"llama": models.LlamaWithConfig, "mistral": models.MistralWithConfig, "qwen2": models.QwenWithConfig,
           "mixtral": models.MixtralWithConfig}[cfg.arch]
    return cls.initialize_model(cf, weights, "bf16", 0)
