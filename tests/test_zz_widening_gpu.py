"""GPU tests of the SURVEY.md section 8f widening rows -- sampling (8f-1) and equal-length request batching (8f-3).  Written after the
round's GPU budget was spent, so the file name sorts after the parity suites: they run last.

fl_forward_sample = fl_forward + the library's LogitsProcessor on row 0.  The kernels are deterministic, so a second cache fed
the same ids produces the same logits; sampling those with the ORACLE's LogitsProcessor (same seed) must give the same ids."""
import numpy as np
import pytest

from oracle import sampling as osamp

from helpers import golden_weights, product_model

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,temperature", [("llama_gqa8", 0.8), ("qwen2", 1.3), ("llama", 0.0)])
def test_forward_sample_equals_oracle_sampler_on_product_logits(name, temperature):
    from fastllm_b200 import models
    cfg, w, g = golden_weights(name)
    model, _ = product_model(cfg, w)
    prompt = np.asarray(g["prompt"], dtype=np.uint32)[None]
    c_logits, c_sample = models.DeviceCache(model.dev, 1, 128), models.DeviceCache(model.dev, 1, 128)
    want_lp, got_lp = osamp.LogitsProcessor(0, temperature), models.LogitsProcessor(0, temperature)
    ids, pos = prompt, 0
    for _ in range(10):
        want = want_lp.sample(c_logits.forward(ids, pos)[0])
        got = c_sample.forward_sample(ids, pos, got_lp)
        assert got == want
        pos += ids.shape[1]
        ids = np.array([[got]], dtype=np.uint32)
    assert c_sample.kv_len() == c_logits.kv_len() == prompt.shape[1] + 9


def test_generate_with_temperature_runs_through_the_adapter():
    """Model.generate(temperature > 0) through the Mistral adapter: ids equal the oracle sampler fed the product's own logits."""
    from fastllm_b200 import models
    cfg, w, g = golden_weights("mistral")
    model, cache = product_model(cfg, w)
    ids, logits = models.Model(model, cache, eos_token_id=None).generate(g["prompt"], 8, temperature=0.7, return_logits=True)
    lp = osamp.LogitsProcessor(0, float(np.float32(0.7)))
    assert ids == [lp.sample(r) for r in logits]


KERNEL_TOL = 3e-3      # tests/test_parity_gpu.py: batch-b decode (dense path) vs batch-1 decode (GEMV / persistent path)


@pytest.mark.parametrize("name", ["mistral", "llama_gqa8"])
def test_generate_batch_tokens_are_the_per_request_argmax(name):
    """generate_batch groups equal-length prompts into [b, t] calls.  Batched and single-sequence decode run different kernels
    (same arithmetic, different summation order), so instead of demanding identical ids the check is teacher-forced: every
    token the batch emitted must be the arg-max of the single-sequence logits up to twice the kernel tolerance."""
    from fastllm_b200 import models
    from oracle import synth
    cfg, w, g = golden_weights(name)
    model, cache = product_model(cfg, w)
    lengths = [9, 5, 9, 9, 5, 12]
    prompts = [list(map(int, synth.token_ids(40 + i, cfg.vocab_size, (n,)))) for i, n in enumerate(lengths)]
    batched = models.Model(model, cache, eos_token_id=None).generate_batch(prompts, 6)
    assert [len(t) for t in batched] == [6] * len(prompts)
    for p, toks in zip(prompts, batched):
        c = model.initialize_cache()
        logits = np.asarray(model.forward(np.asarray([p], dtype=np.uint32), 0, c))[0].reshape(-1)
        pos = len(p)
        for tok in toks:
            assert logits[tok] >= logits.max() - 2 * KERNEL_TOL
            logits = np.asarray(model.forward(np.array([[tok]], dtype=np.uint32), pos, c))[0].reshape(-1)
            pos += 1


def test_embed_many_equals_per_sentence():
    from fastllm_b200 import models
    from oracle import bert as obert
    from oracle import synth
    cfg = obert.BertConfig(128, 4, 2, 256, 64, 1e-12, 200)
    m = models.MiniLMModel(models.BertConfig(cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers, cfg.intermediate_size,
                                             cfg.max_position_embeddings, cfg.layer_norm_eps, cfg.vocab_size),
                           obert.synth_weights(cfg, 5, 0.05))
    sentences = [list(map(int, synth.token_ids(70 + i, cfg.vocab_size, (n,)))) for i, n in enumerate([7, 19, 7, 33, 19, 7, 1])]
    got = m.embed_many(sentences)
    want = obert.MiniLM(cfg, obert.synth_weights(cfg, 5, 0.05))
    for i, s in enumerate(sentences):
        one = m.embed_ids(np.asarray([s], dtype=np.uint32))[0]
        assert float(np.dot(one, got[i])) >= 0.99999 and np.abs(one - got[i]).max() <= 1e-3
        assert float(np.dot(want.embed_ids(np.asarray([s], dtype=np.uint32))[0], got[i])) >= 0.999      # and against the oracle
