"""Generates tests/golden/fulldepth_tinyllama.json: BASELINE.json config 1 at FULL depth on the f32 oracle.

    TinyLlama-1.1B (22 layers, llama.rs:125-145 shape), synthetic weights (oracle/synth.py, seed 0, std 0.02),
    prompt = 128 synthetic ids, 64 greedy steps through the reference's generate loop (models/mod.rs:411-453).

Prompt seed.  Random-init logits over 32000 tokens have top-1/top-2 gaps that are exponentially distributed with mean ~0.28: in
64 steps some gap is almost always < 1e-2, and a step like that says nothing about a kernel (the product's bf16 KV cache alone
moves logits by ~1e-2 against the f32 oracle; SURVEY.md section 7, hard part 1 asks to "choose/record seeds whose margins exceed
the measured error").  `--search N` runs seeds 1..N and prints, per seed, the smallest margin over the first 32 and over all 64
steps; the committed golden uses PROMPT_SEED = the seed with the largest 32-step minimum among 1..300 (ties: larger 64-step
minimum).  Nothing else about the run is selected.

Stored: the 64 greedy ids (+ their SHA-256), the margin of every step, json; and in fulldepth_tinyllama_logits.npz the prompt and
the full f32 logits rows of steps 0 / 31 / 63.  Run here (CPU, ~15 s per seed, 4.4 GB of f32 weights):
    python tests/golden/make_fulldepth_golden.py [--seed S | --search N]
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import causal_lm as ocl   # noqa: E402
from oracle import synth              # noqa: E402

PROBE = [0, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 31999]


PROMPT_SEED = 100
_MODEL = None


def run(seed, kv_dtype="f32", n_prompt=128, n_steps=64):
    global _MODEL
    cfg = ocl.TINYLLAMA
    if _MODEL is None:
        _MODEL = ocl.CausalLM(cfg, ocl.synth_weights(cfg, 0, 0.02), kv_dtype=kv_dtype)
    prompt = synth.token_ids(seed, cfg.vocab_size, (n_prompt,))
    ids, logits = ocl.generate(ocl.make_adapter(_MODEL), prompt, n_steps, eos_id=None, return_logits=True)
    margins = []
    for lg in logits:
        top2 = np.partition(lg, -2)[-2:]
        margins.append(float(top2[1] - top2[0]))
    return prompt, ids, logits, margins


def main():
    if "--search" in sys.argv:
        for seed in range(1, int(sys.argv[sys.argv.index("--search") + 1]) + 1):
            m = run(seed)[3]
            print(seed, round(min(m[:32]), 4), round(min(m), 4), flush=True)
        return
    seed = int(sys.argv[sys.argv.index("--seed") + 1]) if "--seed" in sys.argv else PROMPT_SEED
    prompt, ids, logits, margins = run(seed)
    out = {
        "config": f"TinyLlama-1.1B 22 layers, synth seed 0 std 0.02, prompt synth.token_ids({seed}, 32000, (128,)), 64 greedy steps, f32 oracle",
        "prompt_seed": seed,
        "ids": [int(i) for i in ids],
        "ids_sha256": hashlib.sha256(np.asarray(ids, dtype=np.uint32).tobytes()).hexdigest(),
        "prompt_sha256": hashlib.sha256(np.asarray(prompt, dtype=np.uint32).tobytes()).hexdigest(),
        "margins": margins,
        "min_margin": min(margins),
        "min_margin_first32": min(margins[:32]),
        "probe_index": PROBE,
        "probe_logits": {str(s): [float(logits[s][i]) for i in PROBE] for s in (0, 31, 63)},
    }
    p = os.path.join(ROOT, "tests", "golden", "fulldepth_tinyllama.json")
    json.dump(out, open(p, "w"), indent=1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fulldepth_tinyllama_logits.npz"), prompt=np.asarray(prompt, dtype=np.uint32),
                        logits_0=logits[0], logits_31=logits[31], logits_63=logits[63])
    print("wrote", p, "min margin", out["min_margin"], "sha", out["ids_sha256"])


if __name__ == "__main__":
    main()
