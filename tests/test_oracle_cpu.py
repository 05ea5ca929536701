"""CPU tests: the oracle against the committed golden vectors, the synthetic generator, and host-side rules."""
import numpy as np
import pytest

from oracle import bert as obert
from oracle import causal_lm as ocl
from oracle import mixtral as omix
from oracle import synth

from helpers import GOLDEN, TINY, golden_weights


@pytest.mark.parametrize("name", list(TINY))
def test_oracle_reproduces_golden(name):
    cfg, w, g = golden_weights(name)
    model = ocl.CausalLM(cfg, w)
    for mode, faithful in (("faithful", True), ("poscorrect", False)):
        if f"{mode}_ids" not in g.files:
            continue
        ids, logits = ocl.generate(ocl.make_adapter(model, faithful), g["prompt"], len(g[f"{mode}_ids"]), eos_id=None,
                                   return_logits=True)
        assert ids == list(g[f"{mode}_ids"])
        assert np.abs(np.stack(logits) - g[f"{mode}_logits"]).max() < 1e-5


def test_oracle_bert_golden():
    g = np.load(f"{GOLDEN}/bert_tiny.npz")
    cfg = obert.BertConfig(128, 4, 2, 256, 64, 1e-12, 200)
    w = obert.synth_weights(cfg, int(g["seed"]), float(g["std"]))
    for k in g.files:
        if k.startswith("w:"):
            w[k[2:]] = g[k]
    m = obert.MiniLM(cfg, w)
    assert np.abs(m.forward(g["ids"]) - g["hidden"]).max() < 1e-5
    emb = m.embed_ids(g["ids"])
    assert np.abs(emb - g["embeddings"]).max() < 1e-6
    assert np.allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-5)


def test_router_known_answers():
    g = np.load(f"{GOLDEN}/mixtral_router.npz")
    idx, wts = omix.route_top_k(g["logits"], 2)
    assert (idx == g["idx"]).all()               # exact ties resolve to the LOWER expert index
    assert np.abs(wts - g["wts"]).max() < 1e-7
    assert np.allclose(wts.sum(1), 1.0, atol=1e-6)


def test_adapter_offset_rules():
    """mistral.rs:206-236: offset += 1 per call (NOT += seq_len); cache.rs reset; llama takes pos from the caller."""
    cfg, w, _ = golden_weights("mistral")
    ad = ocl.MistralAdapter(ocl.CausalLM(cfg, w))
    c = ad.initialize_cache()
    ad.forward(np.array([[5, 6, 7, 8]], dtype=np.uint32), 0, c)
    assert c.get_offset() == 1 and ad.model.kv_len == 4
    ad.forward(np.array([[9]], dtype=np.uint32), 4, c)
    assert c.get_offset() == 2 and ad.model.kv_len == 5
    c.reset()
    ad.forward(np.array([[9]], dtype=np.uint32), 0, c)      # offset 0 => KV cleared
    assert ad.model.kv_len == 1


def test_argmax_last_index_wins():
    assert ocl.sample_argmax(np.array([1.0, 3.0, 3.0, 2.0, 3.0, 0.0], dtype=np.float32)) == 4


def test_eos_breaks_before_emit():
    cfg, w, g = golden_weights("llama")
    ids = ocl.generate(ocl.make_adapter(ocl.CausalLM(cfg, w)), g["prompt"], 8, eos_id=int(g["faithful_ids"][2]))
    assert ids == list(g["faithful_ids"][:2])


def test_config_validation_panics():
    bad = ocl.CausalLMConfig("mistral", 100, 64, 32, 1, 3, 3)      # 100 / 3 not integral
    with pytest.raises(AssertionError):
        bad.validate()
    with pytest.raises(AssertionError):
        ocl.CausalLMConfig("mistral", 96, 64, 32, 1, 6, 4).validate()   # 6 % 4 != 0


def test_synth_c_equals_numpy_and_is_normalish():
    a = synth.normal_bf16_bits(3, "model.layers.0.mlp.up_proj.weight", 200_003, use_c=True)
    b = synth.normal_bf16_bits(3, "model.layers.0.mlp.up_proj.weight", 200_003, use_c=False)
    assert (a == b).all()
    x = synth.bf16_bits_to_f32(a)
    assert abs(float(x.mean())) < 5e-4 and abs(float(x.std()) - 0.02) < 3e-4
    assert synth.UNIT_BITS == 0x37DDB3D7
    ids = synth.token_ids(1, 32000, (4, 128))
    assert ids.min() >= 3 and ids.max() < 32000


def test_bf16_rounding_is_rne():
    x = np.array([1.0, 1.00390625, 1.005859375, -2.5, 3.3895314e38], dtype=np.float32)
    import torch
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    assert (synth.round_bf16(x) == want).all()


def test_safetensors_ingest_round_trip(tmp_path):
    """safetensors_io: single-file and sharded (index.json weight_map) checkpoints come back bit-identical through the
    memory-mapped reader, in f32 / f16 / bf16-bits, with the reference's file-selection rule (huggingface.rs:83-121)."""
    import json
    from fastllm_b200 import safetensors_io as sio
    rng = np.random.default_rng(0)
    a = {"model.embed_tokens.weight": rng.standard_normal((7, 5)).astype(np.float32),
         "model.norm.weight": rng.standard_normal((5,)).astype(np.float16),
         "lm_head.weight": rng.integers(0, 65535, (7, 5)).astype(np.uint16)}
    d1 = tmp_path / "single"
    d1.mkdir()
    sio.write_safetensors(str(d1 / "model.safetensors"), a)
    got = {k: np.array(v) for k, v in sio.iter_tensors(str(d1))}
    assert set(got) == set(a) and all(got[k].dtype == a[k].dtype and np.array_equal(got[k], a[k]) for k in a)
    d2 = tmp_path / "sharded"
    d2.mkdir()
    names = list(a)
    sio.write_safetensors(str(d2 / "model-00001-of-00002.safetensors"), {names[0]: a[names[0]]})
    sio.write_safetensors(str(d2 / "model-00002-of-00002.safetensors"), {n: a[n] for n in names[1:]})
    json.dump({"weight_map": {names[0]: "model-00001-of-00002.safetensors", names[1]: "model-00002-of-00002.safetensors",
                              names[2]: "model-00002-of-00002.safetensors"}}, open(d2 / "model.safetensors.index.json", "w"))
    got = {k: np.array(v) for k, v in sio.iter_tensors(str(d2))}
    assert set(got) == set(a) and all(np.array_equal(got[k], a[k]) for k in a)
    with pytest.raises(FileNotFoundError):
        list(sio.iter_tensors(str(tmp_path)))
