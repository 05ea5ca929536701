"""CPU checks of the measurement contract pieces that do not need a GPU: workloads cover BASELINE.json's configs, the committed
ncu capture feeds `roofline.traffic`, the reference arm answers for every workload, the default flags finish within minutes."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_workloads_cover_baseline_configs():
    import bench
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    text = " ".join(w[3] for w in bench.WORKLOADS.values())
    for key in ("TinyLlama", "MiniLM", "Mistral-7B", "Qwen2.5-7B", "Mixtral-8x7B"):
        assert any(key in c for c in base["configs"]) and key in text, key
    assert bench.WORKLOADS["mistral7b_b1"][1:3] == (1, 2048) and bench.WORKLOADS["mistral7b_b64"][1] == 64
    assert bench.WORKLOADS["minilm_256x128"][1:3] == (256, 128) and bench.WORKLOADS["mixtral8x7b_b32"][1] == 32


def test_ncu_traffic_matches_algorithmic_bytes():
    """profiles/r01_persistent_final_full_raw.csv: DRAM bytes per decode step of the persistent kernel == weights once + KV once
    (Mistral-7B b=1, KV 2048: 14.490 GB) within 1 %."""
    import bench
    t = bench.ncu_traffic_per_step()
    assert t is not None and abs(t / 14.490e9 - 1.0) < 0.01, t


def test_reference_arm_answers_for_workloads_without_a_cpu_leg():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "mixtral8x7b_b32"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and "unavailable" in line


def test_peaks_come_from_measured_file_or_documented_fallback():
    import bench
    hbm, src = bench.peaks("hbm")
    tf, _ = bench.peaks("tensor")
    assert 5000 < hbm < 9000 and 1000 < tf < 2500 and ("measured" in src or "fallback" in src)
