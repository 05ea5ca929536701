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
    """profiles/rNN_persistent*_full_raw.csv (newest round): DRAM bytes per decode step of the persistent kernel == weights once +
    KV once (Mistral-7B b=1, KV 2048: 14.490 GB) within 1 %, and the line names the file it comes from."""
    import bench
    t, src = bench.ncu_traffic_per_step()
    assert t is not None and abs(t / 14.490e9 - 1.0) < 0.01, t
    assert "profiles/" in src


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


def test_reference_arm_times_whole_steps_with_every_core_under_a_launcher_that_pins_one_thread():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the CPU arm (rank 0 alone) must still use every host core, and when the
    model fits the host it times WHOLE decode steps (TinyLlama: 4.1 GB f32), so ms_per_step x steps is time actually spent."""
    import time
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    t0 = time.time()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tinyllama_b1", "--gpus", "2",
                        "--steps", "6", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    wall = time.time() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["extrapolated"] is False and line["cpu_baseline"]["extrapolated"] is False
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1) or (os.cpu_count() or 1) == 1
    assert line["steps"] * line["ms_per_step"] / 1e3 < wall       # the claimed steps fit inside the run
    # the other ranks of the launch exit without work and without output
    env["RANK"] = "1"
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tinyllama_b1", "--gpus", "2"],
                        capture_output=True, text=True, timeout=120, env=env)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
