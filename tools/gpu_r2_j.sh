#!/bin/bash
# tcgen05 prefill attention: parity, then A/B on the Qwen 4k prefill
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_attn_streamk_gpu.py -m gpu -q -rA -x --timeout 120 > $O/j_pytest_sk.log 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/j_pytest_sk.log
if [ $rc -ne 0 ]; then tail -40 $O/j_pytest_sk.log; exit 1; fi
timeout 600 python -m pytest tests/test_fulldepth_gpu.py -m gpu -q -rA -x --timeout 300 > $O/j_pytest_fd.log 2>&1; echo "pytest rc=$?" >> $O/j_pytest_fd.log
timeout 300 python tools/survey_perf.py qwen_prefill > $O/j_survey_tc.log 2>&1
FL_ATTN_PREFILL_MMA=1 timeout 300 python tools/survey_perf.py qwen_prefill > $O/j_survey_mma.log 2>&1
tail -3 $O/j_pytest_fd.log; head -12 $O/j_survey_tc.log; head -12 $O/j_survey_mma.log
