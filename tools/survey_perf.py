"""Dev tool: one-shot perf survey of the non-headline configs (per-kernel CUDA-event profile through fl_prof_*).
usage: survey_perf.py [qwen_prefill] [decode8] [decode64] [tinyllama]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fastllm_b200 import models, presets


def show(prof):
    tot = sum(p["ms"] for p in prof)
    for p in sorted(prof, key=lambda p: -p["ms"]):
        print(f"   {p['kernel']:28s} n={p['launches']:5d} avg={1000*p['ms']/p['launches']:9.2f}us share={100*p['ms']/tot:5.1f}% "
              f"{p['bytes']/max(p['ms'],1e-9)/1e6:8.0f} GB/s")
    print(f"   total {tot:.3f} ms")


what = sys.argv[1:] or ["qwen_prefill", "decode8", "decode64", "tinyllama"]
if "qwen_prefill" in what:
    cls, cf = presets.PRESETS["qwen25_7b"]
    model, _ = cls.initialize_model(cf, None, "bf16", 0, random_seed=0)
    T = int(os.environ.get("T", 4096))
    ids = (np.arange(T, dtype=np.uint32) * 7919 % 150000 + 3)[None]
    cache = models.DeviceCache(model.dev, 1, T + 264)
    cache.forward_greedy(ids, 0); cache.reset()
    t0 = time.perf_counter(); cache.forward_greedy(ids, 0); dt = time.perf_counter() - t0
    print(f"qwen2.5-7b prefill T={T}: {dt*1e3:.2f} ms wall ({T/dt:.0f} tok/s; 56.83 TFLOP -> {56.83*T/4096/dt:.0f} TFLOP/s)")
    cache.reset(); models.prof_begin(); cache.forward_greedy(ids, 0); show(models.prof_end())
    first = np.array([5], dtype=np.uint32)
    _, ms = cache.decode_greedy_loop(first, T, 64)
    print(f"qwen2.5-7b decode after prefill: {ms/64:.3f} ms/step {64/ms*1e3:.0f} tok/s")
    del cache, model
for b in (8, 64):
    if f"decode{b}" not in what:
        continue
    cls, cf = presets.PRESETS["mistral7b"]
    model, _ = cls.initialize_model(cf, None, "bf16", 0, random_seed=0)
    ctx = 2048
    cache = models.DeviceCache(model.dev, b, ctx + 80)
    first = np.full((b,), 5, dtype=np.uint32)
    cache.fill_synthetic(b, ctx); cache.decode_greedy_loop(first, ctx, 4)
    cache.fill_synthetic(b, ctx)
    _, ms = cache.decode_greedy_loop(first, ctx, 32)
    step = ms / 32
    bytes_ = model.dev.streamed_bytes() + b * ctx * 131072
    print(f"mistral7b b={b}: {step:.3f} ms/step  {b/step*1e3:.0f} tok/s  {bytes_/step/1e6:.0f} GB/s ({bytes_/step/1e6/6534.1*100:.1f}% of measured HBM)")
    cache.fill_synthetic(b, ctx)
    models.prof_begin(); cache.decode_greedy_loop(first, ctx, 2); show(models.prof_end())
    del cache, model
if "tinyllama" in what:
    cls, cf = presets.PRESETS["tinyllama"]
    model, _ = cls.initialize_model(cf, None, "bf16", 0, random_seed=0)
    ids = (np.arange(128, dtype=np.uint32) * 7919 % 30000 + 3)[None]
    cache = models.DeviceCache(model.dev, 1, 256)
    for _ in range(2):
        cache.reset()
        t0 = time.perf_counter(); nxt = cache.forward_greedy(ids, 0); t1 = time.perf_counter()
        _, ms = cache.decode_greedy_loop(nxt, 128, 64)
        t2 = time.perf_counter()
    print(f"tinyllama C1: prefill128 {1e3*(t1-t0):.2f} ms, 64 decode steps {ms:.2f} ms device ({64/ms*1e3:.0f} tok/s), total wall {1e3*(t2-t0):.2f} ms")
