"""GPU test of the sampling entry point (runs last: the file name sorts after the parity suites).

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
