#!/bin/bash
# round 2, call z: A/B of the consumer-loop rewrite on the three batch-1 models (same box, alternating)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
cp fastllm_b200/libfastllm_b200.so /tmp/lib_new.so
run() { name=$1; wl=$2; timeout 300 python bench.py --workload $wl --steps 128 --warmup 8 --no-cpu --no-extras > $O/z_$name.json 2> $O/z_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/z_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"],1), round(d["ms_per_step"],4))
except Exception as e: print("$name", "ERR", e)
PY
}
for rep in 1 2; do
cp /tmp/lib_new.so fastllm_b200/libfastllm_b200.so
run new_tiny_$rep tinyllama_b1; run new_qwen_$rep qwen25_7b_b1; run new_mistral_$rep mistral7b_b1
cp tools/_build/lib_oldconsumer.so fastllm_b200/libfastllm_b200.so
run old_tiny_$rep tinyllama_b1; run old_qwen_$rep qwen25_7b_b1; run old_mistral_$rep mistral7b_b1
done
cp /tmp/lib_new.so fastllm_b200/libfastllm_b200.so
