"""The reference's own unit tests, re-stated against the product's host-side mirror (fastllm_b200/models.py)
and the C ABI's config failure points.  No GPU: config validation is host logic and runs before the library touches a device.

Reference tests mirrored (paths under /root/reference/src/models):
    cache.rs:49-65, llama.rs:168-205, mistral.rs:255-271, qwen.rs:192-208      cache offset +1 / reset
    cache.rs:68-79 (+ llama/mistral/qwen twins)                                 `Any` downcast -> isinstance here
    llama.rs:208-232, mistral.rs:274-300, qwen.rs:211-237                        config field copy-through (+ the unwrap_or defaults)
    config.rs:61-144, mistral.rs:347-363                                         head-dim / GQA validation (one #[should_panic])
    model_registry.rs / model_initializer.rs:24-27                               get_family / supports_architecture
"""
import ctypes as C

import pytest

from fastllm_b200 import _lib, models


@pytest.mark.parametrize("cls", [models.LlamaCache, models.MistralCache, models.QwenCache])
def test_cache_operations(cls):
    cache = cls()
    assert cache.get_offset() == 0, "Initial offset should be 0"
    cache.increment_offset()
    assert cache.get_offset() == 1, "Offset should be 1 after increment"
    cache.increment_offset()
    assert cache.get_offset() == 2, "Offset should be 2 after second increment"
    cache.reset()
    assert cache.get_offset() == 0, "Offset should be 0 after reset"


def test_cache_as_any():
    for adapter, cls in [(models.LlamaWithConfig, models.LlamaCache), (models.MistralWithConfig, models.MistralCache),
                         (models.QwenWithConfig, models.QwenCache)]:
        cache = adapter.initialize_cache(0, "bf16")
        assert isinstance(cache, cls), f"Should be able to downcast to {cls.__name__}"
        assert not isinstance(cache, str), "Should not be able to downcast to wrong type"
        assert cls.__name__ in repr(cache), "Debug output should contain type name"


def _config_file(**kw):
    base = dict(hidden_size=512, intermediate_size=1024, vocab_size=1000, num_hidden_layers=2, num_attention_heads=8,
                num_key_value_heads=8, rms_norm_eps=1e-5, rope_theta=10000.0, max_position_embeddings=2048)
    base.update(kw)
    return models.ConfigFile(**base)


def test_llama_config_conversion():
    c = models.LlamaWithConfig._to_fl_config(_config_file())
    assert (c.hidden_size, c.intermediate_size, c.vocab_size, c.num_hidden_layers) == (512, 1024, 1000, 2)
    assert (c.num_attention_heads, c.num_key_value_heads, c.max_position_embeddings, c.rope_theta) == (8, 8, 2048, 10000.0)
    assert c.sliding_window == 0 and c.qkv_bias == 0 and c.arch == _lib.FL_ARCH["llama"]
    # llama.rs:31-50 unwrap_or defaults
    d = models.LlamaWithConfig._to_fl_config(_config_file(num_key_value_heads=None, rope_theta=None, max_position_embeddings=None))
    assert (d.num_key_value_heads, d.rope_theta, d.max_position_embeddings) == (8, 10000.0, 4096)


def test_mistral_config_conversion():
    c = models.MistralWithConfig._to_fl_config(_config_file(sliding_window=4096))
    assert (c.hidden_size, c.intermediate_size, c.vocab_size, c.num_hidden_layers) == (512, 1024, 1000, 2)
    assert (c.num_attention_heads, c.num_key_value_heads, c.max_position_embeddings, c.rope_theta) == (8, 8, 2048, 10000.0)
    assert c.sliding_window == 4096 and c.qkv_bias == 0
    # mistral.rs:93-154 defaults: sliding_window Some(unwrap_or(4096)), max_position_embeddings 32768
    d = models.MistralWithConfig._to_fl_config(_config_file(sliding_window=None, max_position_embeddings=None, rope_theta=None))
    assert (d.sliding_window, d.max_position_embeddings, d.rope_theta) == (4096, 32768, 10000.0)
    assert models.MistralWithConfig._to_fl_config(_config_file(sliding_window=5)).sliding_window == 5


def test_qwen_config_conversion():
    c = models.QwenWithConfig._to_fl_config(_config_file(sliding_window=512, rms_norm_eps=1e-6))
    assert (c.hidden_size, c.num_attention_heads, c.num_key_value_heads, c.sliding_window) == (512, 8, 8, 512)
    assert c.qkv_bias == 1 and abs(c.norm_eps - 1e-6) < 1e-12
    # qwen.rs:30-56 defaults
    d = models.QwenWithConfig._to_fl_config(_config_file(num_key_value_heads=None, sliding_window=None, max_position_embeddings=None))
    assert (d.num_key_value_heads, d.sliding_window, d.max_position_embeddings) == (8, 4096, 32768)


def _create(cfg):
    """fl_model_create on a box without a GPU: a bad config must fail on the CONFIG (FL_ERR_INVALID + the reference's message), a good
    one gets as far as the device and fails there (no CPU fallback)."""
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.fl_model_create(C.byref(cfg), C.byref(h))
    msg = lib.fl_last_error().decode()
    if rc == 0:
        lib.fl_model_destroy(h)
    return rc, msg


def test_invalid_head_dim():
    """mistral.rs:347-363 #[should_panic(expected = "hidden_size must be divisible by num_attention_heads")], config.rs:80-99."""
    for adapter in (models.MistralWithConfig, models.QwenWithConfig, models.LlamaWithConfig):
        rc, msg = _create(adapter._to_fl_config(_config_file(hidden_size=500)))
        assert rc == -1 and "hidden_size must be divisible by num_attention_heads" in msg
    rc, msg = _create(models.QwenWithConfig._to_fl_config(_config_file(num_attention_heads=7, num_key_value_heads=7)))
    assert rc == -1 and "hidden_size must be divisible by num_attention_heads" in msg


def test_odd_head_dim():
    """config.rs:36-39: head_dim must be even for RoPE embeddings."""
    rc, msg = _create(models.MistralWithConfig._to_fl_config(_config_file(hidden_size=504, num_attention_heads=8)))   # d = 63
    assert rc == -1 and "head_dim must be even for RoPE embeddings" in msg


def test_invalid_gqa_config():
    """config.rs:123-143: 8 heads, 3 kv heads."""
    rc, msg = _create(models.QwenWithConfig._to_fl_config(_config_file(num_key_value_heads=3)))
    assert rc == -1 and "num_attention_heads must be divisible by num_key_value_heads" in msg


def test_valid_config_reaches_the_device():
    """config.rs:61-78 / 102-120: head_dim 64, GQA 8/4 pass validation; what fails afterwards (here) is the missing device."""
    import torch
    rc, msg = _create(models.QwenWithConfig._to_fl_config(_config_file(num_key_value_heads=4, intermediate_size=2048)))
    if torch.cuda.is_available():
        assert rc == 0
    else:
        assert rc in (-2, -3) and "divisible" not in msg and "head_dim" not in msg


def test_family_and_architecture_registry():
    cases = [(models.LlamaWithConfig, "Llama", "LlamaForCausalLM"), (models.MistralWithConfig, "Mistral", "MistralForCausalLM"),
             (models.QwenWithConfig, "Qwen", "Qwen2ForCausalLM"), (models.MixtralWithConfig, "Mixtral", "MixtralForCausalLM"),
             (models.MiniLMModel, "bert", "BertModel")]
    for cls, family, arch in cases:
        assert cls.get_family() == family
        assert cls.supports_architecture(arch)
        assert not cls.supports_architecture("GPT2LMHeadModel")
    assert not models.MistralWithConfig.supports_architecture("MixtralForCausalLM")      # mistral.rs:244-246
    assert not models.LlamaWithConfig.supports_architecture("MistralForCausalLM")


def test_generate_loop_is_greedy_by_default_and_breaks_on_eos():
    """models/mod.rs:411-453 with the API's default temperature 0.0 (chat.rs:24-25): arg-max, last index wins, EOS never emitted."""
    import numpy as np
    rows = np.full((5, 16), -1.0, dtype=np.float32)
    rows[0, [3, 9]] = 4.0          # tie -> 9
    rows[1, 5] = 1.0
    rows[2, 2] = 7.0               # EOS
    calls = []

    class Adapter:
        def initialize_cache(self, *a):
            return models.LlamaCache()

        def forward(self, ids, pos, cache):
            calls.append((np.asarray(ids).tolist(), pos))
            return rows[len(calls) - 1][None]

    out = models.Model(Adapter(), None, eos_token_id=2).generate([1, 2, 3, 4], 5)
    assert out == [9, 5]
    assert calls == [([[1, 2, 3, 4]], 0), ([[9]], 4), ([[5]], 5)]
