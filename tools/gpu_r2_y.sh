#!/bin/bash
# round 2, call y: full GPU suite + the bench lines of record after the persistent-kernel clean-up and the many-row RMSNorm kernel
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rA --timeout 900 > $O/y_pytest.log 2>&1; echo "pytest rc=$?" >> $O/y_pytest.log
tail -3 $O/y_pytest.log
FL_PREP_ROWS=0 timeout 600 python bench.py --workload qwen25_7b_prefill4k --steps 8 --warmup 3 --no-cpu > $O/y_qwen_prefill_oldprep.json 2> $O/y_qwen_prefill_oldprep.err
timeout 600 python bench.py --workload qwen25_7b_prefill4k --steps 8 --warmup 3 --no-cpu > $O/y_qwen_prefill.json 2> $O/y_qwen_prefill.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/y_bench_default.json 2> $O/y_bench_default.err; echo "rc=$?" >> $O/y_bench_default.err
timeout 300 python __graft_entry__.py smoke > $O/y_smoke.log 2>&1; echo "smoke rc=$?" >> $O/y_smoke.log
timeout 300 python bench.py --workload tinyllama_b1 --steps 64 --warmup 8 --no-cpu > $O/y_tinyllama.json 2> $O/y_tinyllama.err
timeout 300 python bench.py --workload qwen25_7b_b1 --steps 64 --warmup 8 --no-cpu > $O/y_qwen_b1.json 2> $O/y_qwen_b1.err
for f in y_qwen_prefill_oldprep y_qwen_prefill y_bench_default y_tinyllama y_qwen_b1; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["ms_per_step"],4), round(d["roofline"]["frac"],4), round(d["e2e"]["value"],1))
except Exception as e: print("$f", "ERR", e)
PY
done
tail -2 $O/y_smoke.log
