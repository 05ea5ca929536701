"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle and the committed golden vectors.

Tolerances.  The product keeps f32 activations, f32 accumulation and bf16 weights (exactly the oracle's weights);
the ONLY extra rounding against the f32 oracle is the bf16 KV cache (the reference server's model dtype, main.rs:120).
So every case is checked twice:
  * against the oracle run with kv_dtype="bf16" (same cache rounding): logits max-abs <= KERNEL_TOL = 3e-3 (a 1e-7
    summation-order difference, or the 2^-17 residue of the hi/lo bf16 split on the tensor-core path, can flip the bf16
    rounding of a cached K/V element, which then moves logits by ~1e-3; without a flip the measured gap is 1e-6..3e-4)
    -- this pins the kernels themselves; at the BASELINE models' true widths (more cached elements, more flips) the bar is
    KERNEL_TOL_WIDE = 6e-3 (measured 1e-4 .. 3.3e-3);
  * against the pure-f32 oracle / golden vectors: greedy ids IDENTICAL, and logits max-abs <= LOGIT_TOL = 5e-2
    (measured: 0.6e-2 .. 3.2e-2 over 32000 x 5 logits of std 1.3, i.e. ~0.5 % rms -- the bf16 KV rounding, nothing else).
"""
import os

import numpy as np
import pytest

from oracle import causal_lm as ocl
from oracle import synth

from helpers import TINY, golden_weights, product_model

pytestmark = pytest.mark.gpu
LOGIT_TOL = 5e-2
GOLDEN_TOL = 3e-2
KERNEL_TOL = 3e-3
KERNEL_TOL_WIDE = 6e-3


def _generate(model, cache, prompt, n):
    from fastllm_b200 import models
    return models.Model(model, cache, eos_token_id=None).generate(prompt, n, return_logits=True)


@pytest.mark.parametrize("name", ["llama", "llama_gqa8", "mistral", "mistral_sw", "qwen2", "mixtral"])
def test_golden_greedy_and_logits(name):
    """Reference-faithful mode (Mistral/Qwen2: +1-per-call RoPE offset) against tests/golden/causal_*.npz."""
    cfg, w, g = golden_weights(name)
    model, cache = product_model(cfg, w)
    n = len(g["faithful_ids"])
    ids, logits = _generate(model, cache, g["prompt"], n)
    err = float(np.abs(np.stack(logits) - g["faithful_logits"]).max())
    o_ids, o_logits = ocl.generate(ocl.make_adapter(ocl.CausalLM(cfg, w, kv_dtype="bf16")), g["prompt"], n, eos_id=None,
                                   return_logits=True)
    kerr = float(np.abs(np.stack(logits) - np.stack(o_logits)).max())
    print(f"{name}: max-abs logits err vs f32 golden {err:.3e}, vs oracle with bf16 KV {kerr:.3e}")
    assert ids == list(g["faithful_ids"]) == o_ids
    assert kerr <= KERNEL_TOL
    assert err <= GOLDEN_TOL


def test_forward_greedy_and_device_loop_match_host_argmax():
    cfg, w, g = golden_weights("llama_gqa8")
    from fastllm_b200 import models
    model, _ = product_model(cfg, w)
    prompt = np.asarray(g["prompt"], dtype=np.uint32)[None]
    c1 = models.DeviceCache(model.dev, 1, 128)
    logits = c1.forward(prompt, 0)
    c2 = models.DeviceCache(model.dev, 1, 128)
    first = c2.forward_greedy(prompt, 0)
    assert int(first[0]) == models.sample_argmax(logits[0]) == int(g["faithful_ids"][0])
    # device-resident loop: steps 1..7 must reproduce the golden greedy continuation
    out, ms = c2.decode_greedy_loop(first, prompt.shape[1], 7)
    assert list(out[:, 0]) == list(g["faithful_ids"][1:8])
    assert ms > 0 and c2.kv_len() == prompt.shape[1] + 7


def test_graph_replay_equals_eager(monkeypatch):
    cfg, w, g = golden_weights("mistral")
    model, cache = product_model(cfg, w)
    ids_a, logits_a = _generate(model, cache, g["prompt"], 6)
    monkeypatch.setenv("FL_NO_GRAPH", "1")
    monkeypatch.setenv("FL_NO_PDL", "1")
    m2 = model.clone()
    ids_b, logits_b = _generate(m2, m2.initialize_cache(), g["prompt"], 6)
    assert ids_a == ids_b
    assert np.array_equal(np.stack(logits_a), np.stack(logits_b)), "graph+PDL replay must be bit-identical to eager launches"


@pytest.mark.parametrize("name,b", [("llama", 3), ("qwen2", 2)])
def test_batched_decode_equals_per_sequence(name, b):
    """Rows of a batch are independent sequences: batch-b prefill == b separate batch-1 prefills BITWISE (same kernels, row
    results do not depend on their tile neighbours); the batch-b decode step runs on the dense tensor-core path while
    batch-1 decode runs on the GEMV / persistent path, so that comparison is held to the kernel tolerance."""
    cfg, w, _ = golden_weights(name)
    from fastllm_b200 import models
    model, _ = product_model(cfg, w)
    prompts = synth.token_ids(21, cfg.vocab_size, (b, 9))
    cb = models.DeviceCache(model.dev, b, 64)
    lb = [cb.forward(prompts, 0)]
    nxt = np.array([[models.sample_argmax(r)] for r in lb[0]], dtype=np.uint32)
    lb.append(cb.forward(nxt, 9))
    for s in range(b):
        c1 = models.DeviceCache(model.dev, 1, 64)
        l0 = c1.forward(prompts[s:s + 1], 0)
        l1 = c1.forward(nxt[s:s + 1], 9)
        assert np.array_equal(l0[0], lb[0][s])
        assert np.abs(l1[0] - lb[1][s]).max() <= KERNEL_TOL
    # and against the oracle
    want = ocl.CausalLM(cfg, w, kv_dtype="bf16").forward(prompts, 0)
    assert np.abs(want - lb[0]).max() <= KERNEL_TOL


def test_clone_shares_weights_but_not_kv():
    """M: Clone for streaming (mod.rs:155): concurrent streams keep separate KV over shared weights."""
    cfg, w, g = golden_weights("mistral")
    model, cache = product_model(cfg, w)
    twin = model.clone()
    p = np.asarray(g["prompt"], dtype=np.uint32)[None]
    a = model.forward(p, 0, cache)
    cache_t = twin.initialize_cache()
    twin.forward(p[:, :5], 0, cache_t)            # different history in the clone
    b = model.forward(np.array([[7]], dtype=np.uint32), 0, cache)
    model2, cache2 = product_model(cfg, w)
    a2 = model2.forward(p, 0, cache2)
    b2 = model2.forward(np.array([[7]], dtype=np.uint32), 0, cache2)
    assert np.array_equal(a, a2) and np.array_equal(b, b2)


def test_error_behaviour():
    from fastllm_b200 import FastllmError, models
    cfg, w, g = golden_weights("llama")
    model, cache = product_model(cfg, w)
    p = np.asarray(g["prompt"], dtype=np.uint32)[None]
    with pytest.raises(FastllmError):                         # token id out of range
        model.forward(np.array([[cfg.vocab_size]], dtype=np.uint32), 0, cache)
    model.forward(p, 0, cache)
    with pytest.raises(FastllmError):                         # candle Llama: t x t mask on a non-empty cache
        model.forward(p, p.shape[1], cache)
    small = models.DeviceCache(model.dev, 1, 8)
    with pytest.raises(FastllmError):                         # KV capacity
        small.forward(p, 0)
    with pytest.raises(FastllmError):                         # reference: input.dims2()? fails on rank != 2
        model.forward(np.array([1, 2, 3], dtype=np.uint32), 0, cache)
    bad = models.ConfigFile(100, 64, 32, 1, 3, 3)             # mistral.rs:109-127 asserts -> error, never abort
    with pytest.raises(FastllmError):
        models.MistralWithConfig.initialize_model(bad, {}, "bf16", 0)
    with pytest.raises(FastllmError):                         # missing tensors at finalize
        models.LlamaWithConfig.initialize_model(models.ConfigFile(64, 176, 256, 2, 4, 2), {}, "bf16", 0)


def test_device_random_init_matches_host_generator():
    """fl_model_random_init (CUDA) must produce bit-identical bf16 weights to oracle/synth.py: compare through a forward."""
    from fastllm_b200 import models
    cfg = ocl.CausalLMConfig("qwen2", 128, 192, 300, 2, 4, 2, 1e-6, 1e6, 64, 4096, qkv_bias=True)
    cf = models.ConfigFile(128, 192, 300, 2, 4, 2, 1e-6, 1e6, 64)
    dev_model, dev_cache = models.QwenWithConfig.initialize_model(cf, None, "bf16", 0, random_seed=5, std=0.08)
    host_model, host_cache = models.QwenWithConfig.initialize_model(cf, ocl.synth_weights(cfg, 5, 0.08), "bf16", 0)
    p = synth.token_ids(9, 300, (1, 7))
    assert np.array_equal(dev_model.forward(p, 0, dev_cache), host_model.forward(p, 0, host_cache))


@pytest.mark.parametrize("arch", ["tinyllama", "mistral7b", "qwen25_7b"])
def test_true_width_two_layers(arch):
    """True per-layer shapes of the BASELINE models (2 layers; reduced vocab for Qwen to keep the CPU side small):
    prefill + 4 greedy steps against the oracle on the same synthetic weights."""
    from dataclasses import replace
    base = {"tinyllama": ocl.TINYLLAMA, "mistral7b": ocl.MISTRAL_7B, "qwen25_7b": replace(ocl.QWEN25_7B, vocab_size=32064)}[arch]
    cfg = replace(base, num_hidden_layers=2, max_position_embeddings=256)
    w = ocl.synth_weights(cfg, 0, 0.02)
    prompt = synth.token_ids(1, cfg.vocab_size, (24,))
    want_ids, want_logits = ocl.generate(ocl.make_adapter(ocl.CausalLM(cfg, w)), prompt, 4, eos_id=None, return_logits=True)
    _, k_logits = ocl.generate(ocl.make_adapter(ocl.CausalLM(cfg, w, kv_dtype="bf16")), prompt, 4, eos_id=None, return_logits=True)
    model, cache = product_model(cfg, w)
    ids, logits = _generate(model, cache, prompt, 4)
    err = float(np.abs(np.stack(logits) - np.stack(want_logits)).max())
    kerr = float(np.abs(np.stack(logits) - np.stack(k_logits)).max())
    print(f"{arch}: max-abs logits err vs f32 oracle {err:.3e}, vs oracle with bf16 KV {kerr:.3e}")
    assert ids == want_ids and err <= LOGIT_TOL and kerr <= KERNEL_TOL_WIDE


def _tiny_true_width(arch):
    from dataclasses import replace
    base = {"tinyllama": ocl.TINYLLAMA, "mistral7b": ocl.MISTRAL_7B, "qwen25_7b": replace(ocl.QWEN25_7B, vocab_size=32064)}[arch]
    return replace(base, num_hidden_layers=2, max_position_embeddings=256)


@pytest.mark.parametrize("arch", ["tinyllama", "mistral7b", "qwen25_7b"])
def test_persistent_kernel_matches_multikernel_path(arch, monkeypatch):
    """Batch-1 decode runs in the persistent cooperative kernel; the 5-kernels-per-layer path must agree with it
    (different summation order only), and the in-kernel multi-step greedy loop must reproduce step-by-step decoding."""
    from fastllm_b200 import models
    cfg = _tiny_true_width(arch)
    w = ocl.synth_weights(cfg, 0, 0.02)
    prompt = synth.token_ids(1, cfg.vocab_size, (1, 70))           # crosses a 64-token KV page boundary
    model, _ = product_model(cfg, w)
    ca = models.DeviceCache(model.dev, 1, 128)
    first = ca.forward_greedy(prompt, 0)
    ids_loop, _ = ca.decode_greedy_loop(first, 70, 6)                # persistent kernel, 6 steps in one launch
    cb = models.DeviceCache(model.dev, 1, 128)
    tok, step_ids, step_logits = cb.forward_greedy(prompt, 0), [], []
    for s in range(6):                                               # persistent kernel, one launch per step
        lg = cb.forward(tok.reshape(1, 1), 70 + s)
        step_logits.append(lg[0])
        tok = np.array([models.sample_argmax(lg[0])], dtype=np.uint32)
        step_ids.append(int(tok[0]))
    assert list(ids_loop[:, 0]) == step_ids
    monkeypatch.setenv("FL_NO_PERSISTENT", "1")
    cc = models.DeviceCache(model.dev, 1, 128)                       # multi-kernel path
    tok, mk_logits = cc.forward_greedy(prompt, 0), []
    for s in range(6):
        lg = cc.forward(tok.reshape(1, 1), 70 + s)
        mk_logits.append(lg[0])
        tok = np.array([models.sample_argmax(lg[0])], dtype=np.uint32)
        assert int(tok[0]) == step_ids[s]
    err = float(np.abs(np.stack(mk_logits) - np.stack(step_logits)).max())
    print(f"{arch}: persistent vs multi-kernel max-abs logits diff {err:.3e}")
    assert err <= KERNEL_TOL_WIDE


def test_mixtral_router_known_answers_and_batch():
    """Mixtral MoE block: batch rows route independently (batch-3 prefill == 3 single prefills), device synthetic init equals
    host-generated weights, and a true-width single layer (H=4096, I=14336, 8 experts, top-2) matches the oracle."""
    from dataclasses import replace
    from fastllm_b200 import models
    cfg, w, g = golden_weights("mixtral")
    model, _ = product_model(cfg, w)
    prompts = synth.token_ids(5, cfg.vocab_size, (3, 10))
    cb = models.DeviceCache(model.dev, 3, 64)
    lb = cb.forward(prompts, 0)
    want = ocl.CausalLM(cfg, w, kv_dtype="bf16").forward(prompts, 0)
    assert np.abs(lb - want).max() <= KERNEL_TOL
    for s in range(3):
        c1 = models.DeviceCache(model.dev, 1, 64)
        assert np.array_equal(c1.forward(prompts[s:s + 1], 0)[0], lb[s])
    wide = replace(ocl.MIXTRAL_8X7B, num_hidden_layers=1, vocab_size=4096, max_position_embeddings=128)
    ww = ocl.synth_weights(wide, 0, 0.02)
    p = synth.token_ids(1, wide.vocab_size, (1, 9))
    o_ids, o_logits = ocl.generate(ocl.make_adapter(ocl.CausalLM(wide, ww, kv_dtype="bf16")), p[0], 3, eos_id=None, return_logits=True)
    cf = models.ConfigFile(wide.hidden_size, wide.intermediate_size, wide.vocab_size, 1, wide.num_attention_heads, wide.num_key_value_heads,
                           wide.rms_norm_eps, wide.rope_theta, wide.max_position_embeddings, wide.sliding_window, num_local_experts=8,
                           num_experts_per_tok=2)
    m2, c2 = models.MixtralWithConfig.initialize_model(cf, None, "bf16", 0, random_seed=0, std=0.02)
    ids, logits = models.Model(m2, c2, eos_token_id=None).generate(p[0], 3, return_logits=True)
    err = float(np.abs(np.stack(logits) - np.stack(o_logits)).max())
    print(f"mixtral true-width layer: max-abs logits err vs oracle (bf16 KV) {err:.3e}")
    assert ids == o_ids and err <= KERNEL_TOL_WIDE


@pytest.mark.parametrize("name,b,t,steps", [("llama_gqa8", 3, 150, 3), ("qwen2", 4, 131, 3), ("mistral", 5, 70, 2),
                                            ("mistral_sw70", 3, 200, 2)])
def test_tensor_core_attention_long_context(name, b, t, steps):
    """attn_prefill_kernel / attn_gqa_decode_kernel (attn_mma.cuh): prompts spanning several 64-query tiles and 64-token KV pages,
    GQA groups of 8 / 7 / 2 heads, a sliding window that starts inside a later page than the query tile's first key, then
    batched decode steps (one CTA per split x kv head x sequence) -- against the oracle with the same bf16 KV rounding."""
    from dataclasses import replace
    from fastllm_b200 import models
    if name == "mistral_sw70":
        cfg = replace(TINY["mistral"], sliding_window=70)
    else:
        cfg = TINY[name]
    cfg = replace(cfg, max_position_embeddings=512)
    w = ocl.synth_weights(cfg, 11, 0.08)
    model, _ = product_model(cfg, w)
    prompts = synth.token_ids(31, cfg.vocab_size, (b, t))
    oracle = ocl.CausalLM(cfg, w, kv_dtype="bf16")
    cache = models.DeviceCache(model.dev, b, t + steps + 1)
    want, got = oracle.forward(prompts, 0), cache.forward(prompts, 0)
    errs = [float(np.abs(want - got).max())]
    for s in range(steps):
        nxt = np.array([[models.sample_argmax(r)] for r in got], dtype=np.uint32)
        assert [models.sample_argmax(r) for r in want] == [int(x) for x in nxt[:, 0]]
        want, got = oracle.forward(nxt, t + s), cache.forward(nxt, t + s)
        errs.append(float(np.abs(want - got).max()))
    print(f"{name} b={b} t={t}: max-abs logits err per call {['%.2e' % e for e in errs]}")
    assert max(errs) <= KERNEL_TOL


def test_moe_routing_record_equals_the_oracle_router():
    """fl_cache_moe_routing (the diagnostics the sharded-vs-single-GPU checks lean on): the router's picks of a prefill and of a
    decode step equal the oracle's `route_top_k` picks wherever the recorded margin is not a near-tie, in pick order; margins are
    non-negative; a wrong row count is an error."""
    from fastllm_b200 import models
    from fastllm_b200._lib import FastllmError
    from oracle import mixtral as omx
    cfg, w, g = golden_weights("mixtral")
    model, _ = product_model(cfg, w)
    prompts = synth.token_ids(5, cfg.vocab_size, (3, 10))
    oracle = ocl.CausalLM(cfg, w, kv_dtype="bf16")
    cache = models.DeviceCache(model.dev, 3, 64)
    nxt = np.array([[7], [9], [11]], dtype=np.uint32)
    for ids, pos in ((prompts, 0), (nxt, 10)):
        omx.ROUTING_LOG = []
        try:
            oracle.forward(ids, pos)
            want = np.stack(omx.ROUTING_LOG)                       # [L, rows, k]
        finally:
            omx.ROUTING_LOG = None
        cache.forward(ids, pos)
        got, margins = cache.moe_routing(ids.size)
        assert got.shape == want.shape and margins.shape == want.shape[:2]
        assert (margins >= 0).all() and (got >= 0).all() and (got < cfg.num_local_experts).all()
        clear = margins > 1e-3
        assert clear.mean() > 0.9
        assert np.array_equal(got[clear], want[clear])
    with pytest.raises(FastllmError):
        cache.moe_routing(2)


def test_mixtral_masked_prefill_and_grouped_decode_agree(monkeypatch):
    """The MoE block has two execution plans: calls with more than 128 expert rows (prefill) stream every expert over all rows
    with the routing weight as a mask; decode batches gather per-expert row lists and run ONE grouped GEMM pair.  A 150-row
    prefill (masked), batch-3 decode steps (grouped, experts with 0..3 rows), and the same decode steps forced onto the masked
    plan: all against the oracle."""
    from dataclasses import replace
    from fastllm_b200 import models
    cfg = replace(TINY["mixtral"], max_position_embeddings=256)
    w = ocl.synth_weights(cfg, 17, 0.08)
    model, _ = product_model(cfg, w)
    prompts = synth.token_ids(33, cfg.vocab_size, (3, 50))

    def run(masked):
        if masked:
            monkeypatch.setenv("FL_MOE_MASKED", "1")
        else:
            monkeypatch.delenv("FL_MOE_MASKED", raising=False)
        oracle = ocl.CausalLM(cfg, w, kv_dtype="bf16")
        cache = models.DeviceCache(model.dev, 3, 128)
        want, got = oracle.forward(prompts, 0), cache.forward(prompts, 0)
        outs, errs = [got], [float(np.abs(want - got).max())]
        for s in range(4):
            nxt = np.array([[models.sample_argmax(r)] for r in got], dtype=np.uint32)
            want, got = oracle.forward(nxt, 50 + s), cache.forward(nxt, 50 + s)
            outs.append(got)
            errs.append(float(np.abs(want - got).max()))
        return np.stack(outs), max(errs)

    grouped, e_g = run(False)
    masked, e_m = run(True)
    print(f"mixtral tiny: grouped-plan err {e_g:.2e}, masked-plan err {e_m:.2e}, plans differ by {np.abs(grouped - masked).max():.2e}")
    assert e_g <= KERNEL_TOL and e_m <= KERNEL_TOL
    assert np.array_equal(grouped[0], masked[0])          # the prefill runs the masked plan either way

