"""Dev tool: MiniLM-L6 256x128 embeddings/s (device-resident repeats and end-to-end through fl_embed)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fastllm_b200 import models
m = models.MiniLMModel(models.BertConfig(), None, 0, random_seed=0)
b, t = int(os.environ.get("B", 256)), 128
ids = (np.arange(b * t, dtype=np.uint32).reshape(b, t) * 7919 % 30000 + 3).astype(np.uint32)
m.embed_ids(ids)
for _ in range(2):
    _, ms = m.embed_ids_timed(ids, 20)
print(f"device-resident: {ms/20:.3f} ms/batch  -> {b/(ms/20/1e3):.0f} emb/s  ({0.734e12*b/256/(ms/20/1e3)/1e12:.1f} TFLOP/s)")
t0 = time.perf_counter()
for _ in range(20):
    m.embed_ids(ids)
dt = (time.perf_counter() - t0) / 20
print(f"e2e fl_embed:    {dt*1e3:.3f} ms/batch  -> {b/dt:.0f} emb/s")
