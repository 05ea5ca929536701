"""Dev tool: time the persistent decode kernel (FL_MODEL, default Mistral-7B; b=1, KV FL_CTX) under different dev knobs.
usage: pk_sweep.py [static_share_in_32nds[:flags]] ...   (FL_PK_STATIC / FL_PK_FLAGS)"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    from fastllm_b200 import models, presets
    cls, cf = presets.PRESETS[os.environ.get("FL_MODEL", "mistral7b")]
    model, _ = cls.initialize_model(cf, None, "bf16", 0, random_seed=0)
    ctx = int(os.environ.get("FL_CTX", "2048"))
    cache = models.DeviceCache(model.dev, 1, ctx + 200)
    first = np.array([5], dtype=np.uint32)
    cache.fill_synthetic(1, ctx); cache.decode_greedy_loop(first, ctx, 8)
    best = 1e9
    for _ in range(3):
        cache.fill_synthetic(1, ctx)
        _, ms = cache.decode_greedy_loop(first, ctx, 64)
        best = min(best, ms / 64)
    print(f"static={os.environ.get('FL_PK_STATIC')}/32 flags={os.environ.get('FL_PK_FLAGS')}  ms/step={best:.4f}  tok/s={1000/best:.1f}")
else:
    for spec in sys.argv[1:] or ["30"]:
        st, _, flags = spec.partition(":")
        env = dict(os.environ, FL_PK_STATIC=st, FL_PK_FLAGS=flags or "0")
        subprocess.run([sys.executable, __file__, "child"], env=env)
