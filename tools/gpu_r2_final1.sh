#!/bin/bash
# round 2, final 1-GPU lines of record (final build)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rA --timeout 900 > $O/f_pytest.log 2>&1; echo "pytest rc=$?" >> $O/f_pytest.log
tail -2 $O/f_pytest.log
timeout 900 python bench.py > $O/f_bench_default.json 2> $O/f_bench_default.err; echo "rc=$?" >> $O/f_bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/f_bench_reference.json 2> $O/f_bench_reference.err; echo "rc=$?" >> $O/f_bench_reference.err
timeout 600 python bench.py --workload qwen25_7b_prefill4k --steps 8 --warmup 3 --no-cpu > $O/f_qwen_prefill.json 2> $O/f_qwen_prefill.err
timeout 300 python bench.py --workload tinyllama_b1 --steps 64 --warmup 8 --no-cpu > $O/f_tinyllama.json 2> $O/f_tinyllama.err
timeout 300 python bench.py --workload qwen25_7b_b1 --steps 64 --warmup 8 --no-cpu > $O/f_qwen_b1.json 2> $O/f_qwen_b1.err
timeout 300 python bench.py --workload minilm_256x128 --steps 200 --warmup 20 --no-cpu > $O/f_minilm.json 2> $O/f_minilm.err
timeout 400 python bench.py --workload mixtral8x7b_b32 --steps 20 --warmup 5 --no-cpu > $O/f_mixtral.json 2> $O/f_mixtral.err
timeout 300 python __graft_entry__.py smoke > $O/f_smoke.log 2>&1; echo "smoke rc=$?" >> $O/f_smoke.log
timeout 300 python tools/survey_perf.py decode8 decode64 > $O/f_survey.log 2>&1
for f in f_bench_default f_bench_reference f_qwen_prefill f_tinyllama f_qwen_b1 f_minilm f_mixtral; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["ms_per_step"],4), (d.get("roofline") or {}).get("frac"), round(d["e2e"]["value"],1))
except Exception as e: print("$f", "ERR", e)
PY
done
tail -2 $O/f_smoke.log; tail -1 $O/f_bench_default.err
