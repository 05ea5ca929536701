"""Dev tool: does the nvidia-smi clock sampler (bench.py ClockSampler, 50 ms period) perturb the persistent decode kernel?"""
import os, sys, subprocess, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fastllm_b200 import models, presets
cls, cf = presets.PRESETS["mistral7b"]
model, _ = cls.initialize_model(cf, None, "bf16", 0, random_seed=0)
ctx = 2048
cache = models.DeviceCache(model.dev, 1, ctx + 400)
first = np.array([5], dtype=np.uint32)
def run(steps, tag):
    cache.fill_synthetic(1, ctx); cache.decode_greedy_loop(first, ctx, 8)
    res = []
    for _ in range(4):
        cache.fill_synthetic(1, ctx)
        _, ms = cache.decode_greedy_loop(first, ctx, steps)
        res.append(ms / steps)
    print(tag, steps, "steps:", " ".join(f"{r:.4f}" for r in res), flush=True)
run(64, "quiet")
run(256, "quiet")
p = subprocess.Popen(["nvidia-smi", "--query-gpu=timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"], stdout=subprocess.DEVNULL)
time.sleep(1.5)
run(64, "nvidia-smi -lms 50")
run(256, "nvidia-smi -lms 50")
p.terminate(); p.wait()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "500", "-i", "0"], stdout=subprocess.DEVNULL)
time.sleep(1.5)
run(256, "nvidia-smi -lms 500")
p.terminate(); p.wait()
run(256, "quiet again")
