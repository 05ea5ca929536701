#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rA --timeout 900 > $O/k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/k_pytest.log
timeout 300 python tools/survey_perf.py qwen_prefill > $O/k_survey_prefill.log 2>&1
timeout 600 python bench.py --workload qwen25_7b_prefill4k --steps 8 --warmup 3 --no-cpu > $O/k_qwen_prefill.json 2> $O/k_qwen_prefill.err
timeout 300 python bench.py --workload tinyllama_b1 --steps 64 --warmup 8 --no-cpu > $O/k_tinyllama.json 2> $O/k_tinyllama.err
timeout 300 python bench.py --workload qwen25_7b_b1 --steps 64 --warmup 8 --no-cpu > $O/k_qwen_b1.json 2> $O/k_qwen_b1.err
timeout 300 python bench.py --workload mixtral8x7b_b32 --steps 20 --warmup 5 --no-cpu > $O/k_mixtral.json 2> $O/k_mixtral.err
tail -3 $O/k_pytest.log; head -8 $O/k_survey_prefill.log
