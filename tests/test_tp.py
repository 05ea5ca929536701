"""Tensor parallelism: world_size-2 gloo test of the sharding scheme on CPU, and (when 2+ GPUs are visible) TP-2 through
the library with real NCCL against the TP-1 result."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(mode, nproc, out, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "tp_worker.py"), mode, out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return json.load(open(out))


def test_tp_sharding_scheme_gloo_world2(tmp_path):
    """Column-parallel gate/up + row-parallel down with an all-reduce, vocab-parallel head with an all-gather, q/k/v split by
    head: same numbers as the unsharded oracle (numpy over gloo, 2 processes)."""
    assert _run("cpu", 2, str(tmp_path / "cpu.json"), 29631)["ok"]


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_tp2_matches_tp1(tmp_path):
    r1 = _run("gpu", 1, str(tmp_path / "tp1.json"), 29641)
    r2 = _run("gpu", 2, str(tmp_path / "tp2.json"), 29642)
    # expert parallelism (Mixtral): EP-2 == EP-1 == golden greedy ids
    assert r1["mixtral"]["ids"] == r2["mixtral"]["ids"] == r1["mixtral"]["golden_ids"]
    assert np.abs(np.array(r1["mixtral"]["logits"]) - np.array(r2["mixtral"]["logits"])).max() < 3e-3
    # expert parallelism with data-parallel attention + dispatch/combine all-to-all: same logits as one GPU running all 4 sequences
    assert np.abs(np.array(r1["mixtral_dp"]) - np.array(r2["mixtral_dp"])).max() < 3e-3
    # persistent decode kernel with the in-kernel NVLink all-reduce (TP-2) against the single-GPU persistent kernel
    assert r1["wide"]["ids"] == r2["wide"]["ids"] and r1["wide"]["loop_ids"] == r2["wide"]["loop_ids"]
    assert np.abs(np.array(r1["wide"]["logits"]) - np.array(r2["wide"]["logits"])).max() < 6e-3
    # Qwen2 with more ranks than kv heads: replicated kv head + zero-padded query heads
    assert r1["qwen2"]["ids"] == r2["qwen2"]["ids"] == r1["qwen2"]["golden_ids"]
    for k in ("logits", "synth_logits", "batch3"):
        assert np.abs(np.array(r1["qwen2"][k]) - np.array(r2["qwen2"][k])).max() < 3e-3, k
    one, two = r1["mistral"], r2["mistral"]
    assert one["ids"] == two["ids"]
    assert np.abs(np.array(one["logits"]) - np.array(two["logits"])).max() < 3e-3
    assert np.abs(np.array(one["synth_logits"]) - np.array(two["synth_logits"])).max() < 3e-3
    assert np.abs(np.array(one["batch3"]) - np.array(two["batch3"])).max() < 3e-3
