"""Full-depth parity on the BASELINE.json configurations (SURVEY.md section 8c "full TinyLlama 128+64 ids"; north_star: "greedy
token ids must be identical for the first 32 generated tokens, logits within a stated tolerance").

Config 1: TinyLlama-1.1B, ALL 22 layers, a real 128-token prefill and 64 greedy decode steps through the reference's generate
loop (models/mod.rs:411-453), against
  * the committed golden ids of the f32 oracle (tests/golden/fulldepth_tinyllama.json, written by make_fulldepth_golden.py),
  * the f32 oracle re-run on this box (logits error of every step), and
  * the oracle with the product's bf16 KV rounding (pins the kernels: what is left is summation order).

Tolerances at full depth (stated here, measured values are printed by the test):
  LOGIT_TOL_FULL  = 6e-2  max-abs logits vs the pure-f32 oracle over 64 x 32000 logits of std ~1.3 (the bf16 KV cache, 22 layers)
  KERNEL_TOL_FULL = 1.5e-2 max-abs logits vs the oracle with the same bf16 KV rounding
The prompt seed is the one of 1..300 whose f32-oracle top-1/top-2 margin over the first 32 steps is largest (selection rule in
tests/golden/make_fulldepth_golden.py, margins in the golden file): random-init logits have top-2 gaps down to 1e-3, and an id
flip on such a step says nothing about the kernels (hard part 1).  So: the first 32 free-running ids must be
IDENTICAL; past them, and for every teacher-forced step, an arg-max may differ only where the oracle's gap is within 2x the
measured logits error of that step.
"""
import hashlib
import json
import os
import threading

import numpy as np
import pytest

from oracle import causal_lm as ocl
from oracle import synth

from helpers import GOLDEN, TINY, product_model

pytestmark = pytest.mark.gpu
LOGIT_TOL_FULL = 6e-2
KERNEL_TOL_FULL = 1.5e-2


def _golden():
    return json.load(open(os.path.join(GOLDEN, "fulldepth_tinyllama.json")))


def test_tinyllama_full_depth_prompt128_greedy64():
    from fastllm_b200 import models, presets
    g = _golden()
    cfg = ocl.TINYLLAMA
    prompt = synth.token_ids(int(g["prompt_seed"]), cfg.vocab_size, (128,))
    assert hashlib.sha256(prompt.tobytes()).hexdigest() == g["prompt_sha256"]
    cls, cf = presets.PRESETS["tinyllama"]
    # device-side synthetic weights: bit-identical to oracle/synth.py (test_device_random_init_matches_host_generator)
    model, cache = cls.initialize_model(cf, None, "bf16", 0, random_seed=0, std=0.02)
    ids, _ = models.Model(model, cache, eos_token_id=None).generate(prompt, 64, return_logits=True)
    n_same = next((i for i, (a, b) in enumerate(zip(ids, g["ids"])) if a != b), 64)
    print(f"TinyLlama 22L 128+64: first {n_same} of 64 free-running greedy ids identical to the f32 golden "
          f"(oracle margins: min {g['min_margin']:.4f}, first 32 steps {g['min_margin_first32']:.4f})")
    # the device-resident loop (63 steps inside ONE persistent launch) reproduces the step-by-step ids
    c2 = models.DeviceCache(model.dev, 1, 256)
    first = c2.forward_greedy(prompt[None], 0)
    loop_ids, _ = c2.decode_greedy_loop(first, 128, 63)
    assert [int(first[0])] + [int(x) for x in loop_ids[:, 0]] == ids
    # both oracles, re-run on this box's host cores; then the product TEACHER-FORCED with the golden ids so every one of the 64
    # steps is compared on the same history even past a near-tie
    w = ocl.synth_weights(cfg, 0, 0.02)
    o_ids, o_logits = ocl.generate(ocl.make_adapter(ocl.CausalLM(cfg, w)), prompt, 64, eos_id=None, return_logits=True)
    assert o_ids == g["ids"], "the f32 oracle on this box no longer reproduces the committed golden ids"
    kad = ocl.make_adapter(ocl.CausalLM(cfg, w, kv_dtype="bf16"))
    kc = kad.initialize_cache()
    c3 = models.DeviceCache(model.dev, 1, 256)
    lg, klg = c3.forward(prompt[None], 0), kad.forward(prompt[None], 0, kc)
    errs, kerrs, flips = [], [], []
    for s in range(64):
        errs.append(float(np.abs(lg[0] - o_logits[s]).max()))
        kerrs.append(float(np.abs(lg[0] - np.asarray(klg)[0].reshape(-1)).max()))
        if models.sample_argmax(lg[0]) != g["ids"][s]:
            flips.append(s)
            # an id may differ from the f32 oracle's only where the oracle's own top-2 gap is inside twice the measured logits error
            assert g["margins"][s] <= 2 * errs[-1], f"step {s}: arg-max differs although the oracle margin is {g['margins'][s]:.3e}"
        tok = np.array([[g["ids"][s]]], dtype=np.uint32)
        lg, klg = c3.forward(tok, 128 + s), kad.forward(tok, 128 + s, kc)
    err, kerr = max(errs), max(kerrs)
    print(f"TinyLlama 22L 128+64 (teacher-forced): max-abs logits err vs f32 oracle {err:.3e} (tol {LOGIT_TOL_FULL}), vs bf16-KV oracle "
          f"{kerr:.3e} (tol {KERNEL_TOL_FULL}); arg-max == oracle id on {64 - len(flips)}/64 steps, differs at {flips}")
    assert err <= LOGIT_TOL_FULL and kerr <= KERNEL_TOL_FULL
    # north_star: the first 32 generated ids are identical to the reference CPU path's
    assert ids[:32] == g["ids"][:32], f"greedy ids diverge from the f32 oracle at step {n_same}"
    if n_same < 64:
        assert g["margins"][n_same] <= 2 * errs[n_same]
    z = np.load(os.path.join(GOLDEN, "fulldepth_tinyllama_logits.npz"))       # and the committed logits rows (another machine's BLAS)
    assert np.abs(o_logits[31] - z["logits_31"]).max() <= 1e-3


@pytest.mark.parametrize("arch", ["mistral7b", "qwen25_7b"])
def test_true_width_real_vocab_greedy32(arch):
    """True per-layer shapes AND the real vocabulary (Qwen2.5: the 152064-row lm_head), 2 layers, 24-token prompt + 32 greedy
    steps: ids identical to the f32 oracle wherever its margin allows, logits inside the tolerances of test_parity_gpu.py."""
    from dataclasses import replace
    from fastllm_b200 import models
    base = {"mistral7b": ocl.MISTRAL_7B, "qwen25_7b": ocl.QWEN25_7B}[arch]
    cfg = replace(base, num_hidden_layers=2, max_position_embeddings=256)
    w = ocl.synth_weights(cfg, 0, 0.02)
    prompt = synth.token_ids(1, cfg.vocab_size, (24,))
    model, cache = product_model(cfg, w)
    ids, logits = models.Model(model, cache, eos_token_id=None).generate(prompt, 32, return_logits=True)
    # teacher-forced oracles: fed the PRODUCT's ids, so one near-tie cannot derail the comparison of the later steps
    errs, kerrs, margins, flips = [], [], [], 0
    for kv, out in (("f32", errs), ("bf16", kerrs)):
        o = ocl.make_adapter(ocl.CausalLM(cfg, w, kv_dtype=kv))
        c = o.initialize_cache()
        lg = o.forward(prompt[None], 0, c)
        for s in range(32):
            row = np.asarray(lg)[0].reshape(-1)
            out.append(float(np.abs(row - logits[s]).max()))
            if kv == "f32":
                top2 = np.partition(row, -2)[-2:]
                margins.append(float(top2[1] - top2[0]))
                if ocl.sample_argmax(row) != ids[s]:
                    flips += 1
                    assert margins[-1] <= 2 * out[-1], f"step {s}: id differs from the f32 oracle although its margin {margins[-1]:.3e} > 2 x err"
            lg = o.forward(np.array([[ids[s]]], dtype=np.uint32), 24 + s, c)
    print(f"{arch} real vocab: 32 greedy steps, {32 - flips} ids == f32 oracle arg-max (min margin {min(margins):.3e}), "
          f"max-abs logits err {max(errs):.3e} vs f32, {max(kerrs):.3e} vs bf16-KV oracle")
    assert max(errs) <= 5e-2 and max(kerrs) <= 6e-3


def test_smoke_shape_runs_the_persistent_kernel():
    """The shape __graft_entry__.smoke() uses (2 layers, H=2048, d=128) is wide enough for decode_persistent_kernel: its
    profile entry shows up and the multi-kernel path agrees with it."""
    import __graft_entry__ as ge
    from fastllm_b200 import models
    cfg, w, prompt = ge.smoke_case()
    model, _ = product_model(cfg, w)
    c = models.DeviceCache(model.dev, 1, 128)
    tok = c.forward_greedy(prompt[None], 0)
    models.prof_begin()
    c.forward(tok.reshape(1, 1), len(prompt))
    prof = models.prof_end()
    assert [p["kernel"] for p in prof] == ["decode_persistent"], prof


def test_concurrent_streams_on_clones():
    """Streaming requests clone the model handle and run concurrently on shared weights with separate caches, and a stream's
    forward calls may come from different OS threads (models/mod.rs:151-175).  Two threads drive two clones at the same time --
    batch-1 decode in the cooperative persistent kernel, and a batch-3 call on the dense path -- and must reproduce the
    sequential results bit for bit."""
    import __graft_entry__ as ge
    from fastllm_b200 import models
    cfg, w, prompt = ge.smoke_case()
    model, _ = product_model(cfg, w)
    prompts = [prompt, np.roll(prompt, 3)]
    p3 = np.stack([prompt, prompt[::-1], np.roll(prompt, 5)]).astype(np.uint32)

    def job(m, p, out, k):
        c = m.initialize_cache()
        ids, logits = models.Model(m, c, eos_token_id=None).generate(p, 12, return_logits=True)
        c3 = models.DeviceCache(m.dev, 3, 64)
        l3 = [c3.forward(p3, 0)]
        l3.append(c3.forward(np.array([[5], [6], [7]], dtype=np.uint32), p3.shape[1]))
        out[k] = (ids, np.stack(logits), np.stack(l3))

    want = {}
    for k, p in enumerate(prompts):
        job(model.clone(), p, want, k)
    for _ in range(3):
        got, ths = {}, []
        for k, p in enumerate(prompts):
            ths.append(threading.Thread(target=job, args=(model.clone(), p, got, k)))
        for t in ths:
            t.start()
        for t in ths:
            t.join(timeout=300)
            assert not t.is_alive(), "concurrent forward calls on two clones did not finish"
        for k in range(2):
            assert got[k][0] == want[k][0]
            assert np.array_equal(got[k][1], want[k][1]) and np.array_equal(got[k][2], want[k][2])


def test_extra_tensors_in_the_checkpoint_are_ignored(tmp_path):
    """load_model hands EVERY tensor of the checkpoint to initialize_model and VarBuilder ignores what the architecture does not
    read (huggingface.rs:81-135): rotary inv_freq buffers of older Llama exports, `visual.*` of the Qwen2.5-VL checkpoints the
    Qwen adapter advertises (qwen.rs:178-183).  Through a real safetensors file."""
    from safetensors.numpy import save_file
    from fastllm_b200 import models, safetensors_io
    cfg = TINY["qwen2"]
    w = ocl.synth_weights(cfg, 3, 0.08)
    extra = dict(w)
    extra["model.layers.0.self_attn.rotary_emb.inv_freq"] = np.arange(8, dtype=np.float32)
    extra["visual.blocks.0.attn.qkv.weight"] = np.ones((4, 4), dtype=np.float32)
    extra["lm_head.bias_not_a_thing"] = np.zeros((3,), dtype=np.float32)
    save_file(extra, str(tmp_path / "model.safetensors"))
    cf = models.ConfigFile(cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.num_hidden_layers, cfg.num_attention_heads,
                           cfg.num_key_value_heads, cfg.rms_norm_eps, cfg.rope_theta, cfg.max_position_embeddings, cfg.sliding_window)
    m_file, c_file = models.QwenWithConfig.initialize_model(cf, safetensors_io.iter_tensors(str(tmp_path)), "bf16", 0)
    m_dict, c_dict = product_model(cfg, w)
    p = synth.token_ids(4, cfg.vocab_size, (1, 9))
    assert np.array_equal(m_file.forward(p, 0, c_file), m_dict.forward(p, 0, c_dict))


@pytest.mark.parametrize("no_dense", [False, True])
def test_sliding_window_multi_token_call_on_a_non_empty_cache(no_dense, monkeypatch):
    """candle's Mistral/Qwen2 mask is [t, t] over the NEW tokens with zeros concatenated for the cached columns: cached keys stay
    visible, the window thins out the new tokens only (oracle/causal_lm.py _mask).  A second multi-token call with t > sw + 1 on a
    non-empty cache, then a decode step -- on the dense path (attn_prefill_kernel) and, with FL_NO_DENSE=1, on the row-per-CTA
    path (attn_decode_kernel)."""
    from dataclasses import replace
    from fastllm_b200 import models
    if no_dense:
        monkeypatch.setenv("FL_NO_DENSE", "1")
    cfg = replace(TINY["mistral"], sliding_window=5, max_position_embeddings=512)
    w = ocl.synth_weights(cfg, 12, 0.08)
    model, _ = product_model(cfg, w)
    for b, t0, t1 in [(1, 70, 12), (3, 9, 80)]:
        oracle = ocl.CausalLM(cfg, w, kv_dtype="bf16")
        cache = models.DeviceCache(model.dev, b, 256)
        a = synth.token_ids(50 + b, cfg.vocab_size, (b, t0))
        nxt = synth.token_ids(60 + b, cfg.vocab_size, (b, t1))
        errs = [float(np.abs(oracle.forward(a, 0) - cache.forward(a, 0)).max())]
        errs.append(float(np.abs(oracle.forward(nxt, 1) - cache.forward(nxt, 1)).max()))      # adapter rule: offset + 1 per call
        one = synth.token_ids(70 + b, cfg.vocab_size, (b, 1))
        errs.append(float(np.abs(oracle.forward(one, 2) - cache.forward(one, 2)).max()))
        print(f"sliding window 5, b={b}: prefill {t0}, then {t1} tokens on the non-empty cache, then 1: max-abs errs {['%.2e' % e for e in errs]}")
        assert max(errs) <= 3e-3


def test_batched_decode_graph_captured_at_a_short_context_stays_correct():
    """The batched-decode step is ONE CUDA graph per batch size, captured at the first decode step and replayed as the context
    grows: its attention grid must not depend on the KV length at capture time.  Capture after a 4-token prompt, then decode
    across three 64-token pages, logits against the oracle (teacher-forced with the product's ids)."""
    from fastllm_b200 import models
    cfg = TINY["llama_gqa8"]
    from dataclasses import replace
    cfg = replace(cfg, max_position_embeddings=512)
    w = ocl.synth_weights(cfg, 13, 0.08)
    model, _ = product_model(cfg, w)
    b = 3
    oracle = ocl.CausalLM(cfg, w, kv_dtype="bf16")
    cache = models.DeviceCache(model.dev, b, 256)
    ids = synth.token_ids(80, cfg.vocab_size, (b, 4))
    want, got = oracle.forward(ids, 0), cache.forward(ids, 0)
    worst = float(np.abs(want - got).max())
    for s in range(150):
        nxt = np.array([[models.sample_argmax(r)] for r in got], dtype=np.uint32)
        want, got = oracle.forward(nxt, 4 + s), cache.forward(nxt, 4 + s)
        worst = max(worst, float(np.abs(want - got).max()))
    print(f"batch-3 decode over 154 tokens with the graph captured at 4: max-abs logits err {worst:.2e}")
    # Tolerance of a LONG dense-path decode: the tensor-core path carries activations as hi + lo bf16 pairs (2^-17 residue), which
    # flips the bf16 rounding of a freshly cached K / V element ~100x more often than the f32 GEMV path's summation-order noise;
    # every flip moves later logits by ~1e-3 and this config shares ONE kv head between 8 query heads.  Measured over 150 steps:
    # 1e-3 .. 8e-3, not growing with the step index; the bar is the north_star's 1e-2 (tools/diag_batch3.py; the batch-1 path against the same oracle: 3e-4 .. 1e-3).
    assert worst <= 1e-2


def test_minilm_baseline_batch_256x128():
    """BASELINE.json config 2 at its full size: all-MiniLM-L6-v2 shapes, 256 sentences x 128 tokens, against the f32 oracle
    (north_star: cosine >= 0.999 per embedding; max-abs on the unit-norm components reported and held to 2e-2)."""
    from oracle import bert as obert
    from fastllm_b200 import models
    cfg = obert.MINILM_L6
    w = obert.synth_weights(cfg, 0, 0.02)
    ids = synth.token_ids(2, cfg.vocab_size, (256, 128))
    want = obert.MiniLM(cfg, w).embed_ids(ids)
    m = models.MiniLMModel(models.BertConfig(), None, 0, random_seed=0, std=0.02)
    got = m.embed_ids(ids)
    cos = (got * want).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(want, axis=1))
    print(f"MiniLM-L6 256x128: min cosine {cos.min():.6f}, max-abs {np.abs(got - want).max():.3e}")
    assert cos.min() >= 0.999 and np.abs(got - want).max() <= 2e-2
