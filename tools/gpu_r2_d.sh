#!/bin/bash
# 1-GPU perf A/B: persistent-kernel micro-opts, hi|lo single-MMA decode GEMM, attention wave knob
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fulldepth_gpu.py -m gpu -q -rA --timeout 600 -x > $O/d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/d_pytest.log
timeout 300 python bench.py --steps 64 --warmup 8 --no-cpu --no-extras > $O/d_b1.json 2> $O/d_b1.err
timeout 300 python tools/survey_perf.py decode8 decode64 > $O/d_survey.log 2>&1
for wv in 3 6 8 12; do
  FL_ATTN_WAVES=$wv timeout 300 python tools/survey_perf.py decode64 > $O/d_survey_waves$wv.log 2>&1
done
timeout 400 python bench.py --workload mixtral8x7b_b32 --steps 32 --warmup 4 --no-cpu > $O/d_mixtral.json 2> $O/d_mixtral.err
timeout 300 python bench.py --workload tinyllama_b1 --steps 64 --warmup 8 --no-cpu > $O/d_tinyllama.json 2> $O/d_tinyllama.err
timeout 300 python bench.py --workload qwen25_7b_b1 --steps 64 --warmup 8 --no-cpu > $O/d_qwen_b1.json 2> $O/d_qwen_b1.err
