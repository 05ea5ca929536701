#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_fulldepth_gpu.py tests/test_parity_gpu.py -m gpu -q -x --timeout 300 > $O/z2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/z2_pytest.log
tail -3 $O/z2_pytest.log
for wl in tinyllama_b1 qwen25_7b_b1 mistral7b_b1; do timeout 300 python bench.py --workload $wl --steps 128 --warmup 8 --no-cpu --no-extras > $O/z2_$wl.json 2> $O/z2_$wl.err; python - <<PY
import json
try:
    d=json.loads(open("$O/z2_$wl.json").read().strip().splitlines()[-1]); print("$wl", round(d["value"],1), round(d["ms_per_step"],4), round(d["roofline"]["frac"],4))
except Exception as e: print("$wl", "ERR", e)
PY
done
