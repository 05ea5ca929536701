#!/bin/bash
# stream-K decode attention: parity first, then A/B against the per-(split, kv head, sequence) kernel
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 240 python -m pytest tests/test_attn_streamk_gpu.py -m gpu -q -rA -x --timeout 200 > $O/i_pytest_sk.log 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/i_pytest_sk.log
if [ $rc -ne 0 ]; then tail -30 $O/i_pytest_sk.log; exit 1; fi
timeout 900 python -m pytest tests/test_zz_widening_gpu.py tests/test_parity_gpu.py -m gpu -q -rA -x --timeout 600 > $O/i_pytest.log 2>&1; echo "pytest rc=$?" >> $O/i_pytest.log
timeout 300 python tools/survey_perf.py decode8 decode64 > $O/i_survey_new.log 2>&1
FL_ATTN_SK_STAGES=2 timeout 300 python tools/survey_perf.py decode8 decode64 > $O/i_survey_st2.log 2>&1
FL_ATTN_OLD=1 timeout 300 python tools/survey_perf.py decode8 > $O/i_survey_old.log 2>&1
timeout 300 python bench.py --workload mixtral8x7b_b32 --steps 20 --warmup 5 --no-cpu > $O/i_mixtral.json 2> $O/i_mixtral.err
FL_ATTN_SK_STAGES=2 timeout 300 python bench.py --workload mixtral8x7b_b32 --steps 20 --warmup 5 --no-cpu > $O/i_mixtral_st2.json 2> $O/i_mixtral_st2.err
tail -3 $O/i_pytest.log; for f in new st2 old; do grep -E "mistral7b|attn" $O/i_survey_$f.log; done
