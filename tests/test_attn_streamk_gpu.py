"""Batched-decode attention (attn_sk_decode_kernel, csrc/attn_mma.cuh): a fixed grid of resident CTAs cuts the step's whole K|V page
stream into equal ranges, so a (sequence, kv head) pair may lie inside one CTA's range, start in one and end in the next, or span
many.  These cases need MORE pages than resident CTAs (296-444), which the tiny parity configs never reach: here 96-192 ragged
sequences of 1 .. 15 pages decode together, at head dims 32 (unswizzled box), 64 and 128 (one / two swizzled TMA boxes per page).
The prompts (up to 960 tokens, prefilled on empty slots) also run the tcgen05 prefill attention (attn_prefill_tc_kernel: head dims 64 /
128, several 128-query tiles, a 100-token sliding window in one case).  Every row against a per-sequence oracle with the same bf16 KV rounding (8f-3)."""
import numpy as np
import pytest

from oracle import causal_lm as ocl
from oracle import synth

from helpers import product_model

pytestmark = pytest.mark.gpu

CASES = {
    # name: (config, sequences, lengths cycled over the sequences)
    "d32_gqa2": (ocl.CausalLMConfig("mistral", 128, 224, 320, 2, 4, 2, 1e-5, 1e4, 1024, 4096), 120, [20, 64, 65, 700, 130, 900, 1, 333]),
    "d64_gqa2": (ocl.CausalLMConfig("llama", 256, 352, 384, 2, 4, 2, 1e-5, 1e4, 1024), 96, [900, 3, 64, 129, 500, 65]),
    "d64_window": (ocl.CausalLMConfig("mistral", 256, 352, 384, 2, 4, 2, 1e-5, 1e4, 1024, 100), 96, [900, 3, 64, 129, 500, 65]),
    "d128_gqa4": (ocl.CausalLMConfig("qwen2", 512, 704, 384, 2, 4, 1, 1e-6, 1e6, 1024, 4096, qkv_bias=True), 192, [64, 777, 5, 130, 960, 33]),
}


@pytest.mark.parametrize("name", list(CASES) + ["d128_gqa4/tau0", "d64_window/mma"])
def test_stream_k_decode_attention_over_more_pages_than_ctas(name, monkeypatch):
    from fastllm_b200 import models
    # the tcgen05 prefill kernel moves a row's exponent offset lazily (O in TMEM is rescaled only on a jump of > 2^8, which random
    # weights never produce): FL_ATTN_TC_TAU=0 rescales on every new row max, so that path is covered too; FL_ATTN_PREFILL_MMA=1
    # keeps the mma.sync prefill kernel (the path of calls on non-empty caches) on the same inputs
    name, _, variant = name.partition("/")
    if variant == "tau0":
        monkeypatch.setenv("FL_ATTN_TC_TAU", "0")
    if variant == "mma":
        monkeypatch.setenv("FL_ATTN_PREFILL_MMA", "1")
    cfg, nseq, cycle = CASES[name]
    w = ocl.synth_weights(cfg, 29, 0.06)
    model, _ = product_model(cfg, w)
    cache = models.DeviceCache(model.dev, nseq, 1024)
    per_call = cfg.arch != "llama"                     # Mistral / Qwen2 adapters: RoPE offset + 1 per call
    lengths = [cycle[i % len(cycle)] for i in range(nseq)]
    pages = sum((n + 1 + 63) // 64 for n in lengths) * cfg.num_key_value_heads
    assert pages > 2 * 444                             # several pages per resident CTA
    oracles, toks, pos, errs = [], [], [], []
    for i, n in enumerate(lengths):
        prompt = synth.token_ids(300 + i, cfg.vocab_size, (n,))
        o = ocl.CausalLM(cfg, w, kv_dtype="bf16")
        want = o.forward(prompt[None], 0)[0]
        got = cache.forward_slots([i], prompt[None], [0])[0]
        errs.append(float(np.abs(got - want).max()))
        oracles.append(o)
        toks.append(models.sample_argmax(want))        # teacher-forced with the oracle's pick
        pos.append(1 if per_call else n)
    prefill_err = max(errs)
    step_errs = []
    for _ in range(3):
        ids = np.array([[t] for t in toks], dtype=np.uint32)
        got = cache.forward_slots(list(range(nseq)), ids, pos)
        for i in range(nseq):
            want = oracles[i].forward(ids[i:i + 1], pos[i])[0]
            step_errs.append(float(np.abs(got[i] - want).max()))
            toks[i] = models.sample_argmax(want)
            pos[i] += 1
    print(f"{name}: {nseq} ragged sequences, {pages} K|V pages per layer: prefill err {prefill_err:.2e}, decode err {max(step_errs):.2e}")
    # long dense-path decode tolerance (tests/test_zz_widening_gpu.py): bf16 rounding flips of cached K / V under the hi + lo split
    assert prefill_err <= 1e-2 and max(step_errs) <= 1e-2
