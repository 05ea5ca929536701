#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rA --timeout 900 > $O/g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/g_pytest.log
timeout 300 python bench.py --workload minilm_256x128 --steps 30 --warmup 5 --no-cpu > $O/g_minilm.json 2> $O/g_minilm.err
FL_BERT_LNFUSE=1 timeout 300 python bench.py --workload minilm_256x128 --steps 30 --warmup 5 --no-cpu > $O/g_minilm_lnfuse.json 2> $O/g_minilm_lnfuse.err
FL_BERT_BRES=1 timeout 300 python bench.py --workload minilm_256x128 --steps 30 --warmup 5 --no-cpu > $O/g_minilm_bres.json 2> $O/g_minilm_bres.err
timeout 300 python tools/survey_perf.py decode8 decode64 > $O/g_survey.log 2>&1
FL_FUSE=1 timeout 300 python tools/survey_perf.py decode8 decode64 > $O/g_survey_fuse.log 2>&1
FL_FUSE=1 timeout 400 python bench.py --workload mixtral8x7b_b32 --steps 32 --warmup 4 --no-cpu > $O/g_mixtral_fuse.json 2> $O/g_mixtral_fuse.err
timeout 400 python bench.py --workload mixtral8x7b_b32 --steps 32 --warmup 4 --no-cpu > $O/g_mixtral.json 2> $O/g_mixtral.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file $O/g_launches_minilm.csv \
    python bench.py --workload minilm_256x128 --steps 4 --warmup 3 --no-cpu > $O/g_ncu_minilm.log 2>&1
