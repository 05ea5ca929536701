"""Per-step error trace of a batch-3 decode whose CUDA graph was captured at a 4-token context (tests/test_fulldepth_gpu.py)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from dataclasses import replace
from oracle import causal_lm as ocl, synth
from helpers import TINY, product_model
from fastllm_b200 import models
import __graft_entry__ as g
g.build()
cfg = replace(TINY["llama_gqa8"], max_position_embeddings=512)
w = ocl.synth_weights(cfg, 13, 0.08)
model, _ = product_model(cfg, w)
b = 3
oracle = ocl.CausalLM(cfg, w, kv_dtype="bf16")
of32 = ocl.CausalLM(cfg, w)
cache = models.DeviceCache(model.dev, b, 256)
singles = [models.DeviceCache(model.dev, 1, 256) for _ in range(b)]
ids = synth.token_ids(80, cfg.vocab_size, (b, 4))
want, got = oracle.forward(ids, 0), cache.forward(ids, 0)
w32 = of32.forward(ids, 0)
one = np.concatenate([singles[i].forward(ids[i:i + 1], 0) for i in range(b)])
print("logits std", float(np.std(want)))
for s in range(150):
    nxt = np.array([[models.sample_argmax(r)] for r in got], dtype=np.uint32)
    want, got = oracle.forward(nxt, 4 + s), cache.forward(nxt, 4 + s)
    w32 = of32.forward(nxt, 4 + s)
    one = np.concatenate([singles[i].forward(nxt[i:i + 1], 4 + s) for i in range(b)])
    if s % 10 == 9 or s > 140:
        print(s, "dense-vs-bf16kv %.2e  b1path-vs-bf16kv %.2e  dense-vs-b1path %.2e  bf16kv-vs-f32 %.2e" % (
            np.abs(want - got).max(), np.abs(want - one).max(), np.abs(got - one).max(), np.abs(want - w32).max()))
