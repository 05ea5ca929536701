#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rA --timeout 900 > $O/e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/e_pytest.log
timeout 300 python tools/survey_perf.py decode8 decode64 > $O/e_survey.log 2>&1
for wv in 2 4; do
  FL_ATTN_WAVES=$wv timeout 300 python tools/survey_perf.py decode64 > $O/e_survey_waves$wv.log 2>&1
done
timeout 400 python bench.py --workload qwen25_7b_prefill4k --steps 8 --warmup 3 > $O/e_qwen_prefill.json 2> $O/e_qwen_prefill.err
FL_PREFILL_BN128=1 timeout 400 python bench.py --workload qwen25_7b_prefill4k --steps 8 --warmup 3 > $O/e_qwen_prefill_bn128.json 2> $O/e_qwen_prefill_bn128.err
timeout 300 python tools/survey_perf.py qwen_prefill > $O/e_survey_prefill.log 2>&1
