"""Dev tool: dense (tcgen05) path timings: Mistral-7B batched decode at KV 2048 and TinyLlama 128-token prefill."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fastllm_b200 import models, presets
which = sys.argv[1] if len(sys.argv) > 1 else "decode"
if which == "decode":
    cls, cf = presets.PRESETS["mistral7b"]
    model, _ = cls.initialize_model(cf, None, "bf16", 0, random_seed=0)
    for b in [int(x) for x in (sys.argv[2:] or ["8", "64"])]:
        ctx = 2048
        cache = models.DeviceCache(model.dev, b, ctx + 80)
        first = np.full((b,), 5, dtype=np.uint32)
        cache.fill_synthetic(b, ctx); cache.decode_greedy_loop(first, ctx, 4)
        cache.fill_synthetic(b, ctx)
        _, ms = cache.decode_greedy_loop(first, ctx, 32)
        step = ms / 32
        bytes_ = model.dev.streamed_bytes() + b * ctx * 131072
        print(f"mistral7b b={b}: {step:.3f} ms/step  {b/step*1e3:.0f} tok/s  {bytes_/step/1e6:.0f} GB/s ({bytes_/step/1e6/6534.1*100:.1f}% of measured HBM)")
        if os.environ.get("PROF"):
            cache.fill_synthetic(b, ctx)
            models.prof_begin(); cache.decode_greedy_loop(first, ctx, 2); prof = models.prof_end()
            tot = sum(p["ms"] for p in prof)
            for p in prof: print(f"   {p['kernel']:26s} n={p['launches']:4d} avg={1000*p['ms']/p['launches']:8.2f}us share={100*p['ms']/tot:5.1f}% {p['bytes']/max(p['ms'],1e-9)/1e6:8.0f} GB/s")
        del cache
else:
    cls, cf = presets.PRESETS["tinyllama"]
    model, _ = cls.initialize_model(cf, None, "bf16", 0, random_seed=0)
    for T in (128, 512):
        ids = (np.arange(T, dtype=np.uint32) * 7919 % 30000 + 3)[None]
        cache = models.DeviceCache(model.dev, 1, T + 8)
        cache.forward(ids, 0); cache.reset()
        t0 = time.perf_counter(); cache.forward(ids, 0); dt = time.perf_counter() - t0
        print(f"tinyllama prefill T={T}: {dt*1e3:.2f} ms  ({T/dt:.0f} tok/s)")
