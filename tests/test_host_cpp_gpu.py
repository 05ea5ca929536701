"""The compiled (C++) host mirror of the reference's trait API must reproduce the oracle's greedy ids."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import causal_lm as ocl

from helpers import golden_weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "host", "_build", "host_selftest")


def _write_inputs(tmp, cfg, w):
    names = [n for n, _, _ in ocl.tensor_names(cfg)]
    with open(os.path.join(tmp, "manifest.txt"), "w") as mf, open(os.path.join(tmp, "weights.bin"), "wb") as wf:
        mf.write(f"{cfg.hidden_size} {cfg.intermediate_size} {cfg.vocab_size} {cfg.num_hidden_layers} {cfg.num_attention_heads} "
                 f"{cfg.num_key_value_heads} {cfg.rms_norm_eps} {cfg.rope_theta} {cfg.max_position_embeddings}\n")
        for n in names:
            a = np.ascontiguousarray(w[n], dtype=np.float32)
            mf.write(f"{n} {a.ndim} {' '.join(str(d) for d in a.shape)}\n")
            wf.write(a.tobytes())


def test_host_selftest_binary_exists():
    assert os.path.exists(EXE), "host/_build/host_selftest not built by __graft_entry__.build()"


@pytest.mark.gpu
def test_cpp_host_mirror_generates_oracle_ids(tmp_path):
    cfg, w, g = golden_weights("llama_gqa8")
    _write_inputs(str(tmp_path), cfg, w)
    n = len(g["faithful_ids"])
    out = subprocess.run([EXE, str(tmp_path / "manifest.txt"), str(tmp_path / "weights.bin"), str(n)] + [str(int(t)) for t in g["prompt"]],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert json.loads(out.stdout.strip().splitlines()[0]) == [int(t) for t in g["faithful_ids"]]
    assert "expected error" in out.stderr          # Llama t x t mask failure point is an error, not a crash


@pytest.mark.gpu
def test_cpp_continuous_batcher_equals_the_python_mirror(tmp_path):
    """host/fastllm_host.hpp ContinuousBatcher (compiled) and fastllm_b200/models.py ContinuousBatcher issue the same fl_forward_slots
    calls in the same order on the same weights: the ids must be identical (5 requests of different lengths over 2 slots)."""
    from fastllm_b200 import models
    from oracle import synth
    from helpers import product_model
    cfg, w, g = golden_weights("llama_gqa8")
    _write_inputs(str(tmp_path), cfg, w)
    prompts = [list(map(int, synth.token_ids(200 + i, cfg.vocab_size, (n,)))) for i, n in enumerate([7, 3, 19, 5, 11])]
    args = []
    for i, p in enumerate(prompts):
        args += (["/"] if i else []) + [str(t) for t in p]
    out = subprocess.run([EXE, str(tmp_path / "manifest.txt"), str(tmp_path / "weights.bin"), "5", "--batch", "2"] + args,
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    model, _ = product_model(cfg, w)
    want = models.ContinuousBatcher(model, max_batch=2, eos_token_id=None).generate(prompts, 5)
    assert json.loads(out.stdout.strip().splitlines()[0]) == want
