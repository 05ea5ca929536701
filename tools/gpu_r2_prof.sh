#!/bin/bash
# round-2 profiler pass (1 GPU).  Every ncu run follows a plain run of the same command that exited 0.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
set -x
# (a) MiniLM: plain, A/B knobs, then launch list + full capture of the two new GEMM kernels
timeout 300 python bench.py --workload minilm_256x128 --steps 30 --warmup 5 --no-cpu > $O/p_minilm.json 2> $O/p_minilm.err || exit 1
FL_BERT_NO_BRES=1 timeout 300 python bench.py --workload minilm_256x128 --steps 30 --warmup 5 --no-cpu > $O/p_minilm_nobres.json 2> $O/p_minilm_nobres.err
FL_BERT_NO_LNFUSE=1 timeout 300 python bench.py --workload minilm_256x128 --steps 30 --warmup 5 --no-cpu > $O/p_minilm_nolnfuse.json 2> $O/p_minilm_nolnfuse.err
FL_BERT_NO_BRES=1 FL_BERT_NO_LNFUSE=1 timeout 300 python bench.py --workload minilm_256x128 --steps 30 --warmup 5 --no-cpu > $O/p_minilm_r1plan.json 2> $O/p_minilm_r1plan.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file $O/p_launches_minilm.csv \
    python bench.py --workload minilm_256x128 --steps 4 --warmup 3 --no-cpu > $O/p_ncu_minilm.log 2>&1
FL_BERT_NO_BRES=1 FL_BERT_NO_LNFUSE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file $O/p_launches_minilm_r1plan.csv \
    python bench.py --workload minilm_256x128 --steps 4 --warmup 3 --no-cpu > $O/p_ncu_minilm_r1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bert_gemm -s 8 -c 4 -o $O/p_bert_gemm_full -f \
    python bench.py --workload minilm_256x128 --steps 4 --warmup 3 --no-cpu > $O/p_ncu_bert_full.log 2>&1
ncu -i $O/p_bert_gemm_full.ncu-rep --page raw --csv > $O/p_bert_gemm_full_raw.csv 2>/dev/null
# (b) default bench: plain, launch list, full capture of the persistent decode kernel
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-extras > $O/p_default.json 2> $O/p_default.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/p_launches_default.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu --no-extras > $O/p_ncu_default.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_persistent -s 2 -c 1 -o $O/p_persistent_full -f \
    python bench.py --steps 8 --warmup 3 --no-cpu --no-extras > $O/p_ncu_persistent.log 2>&1
ncu -i $O/p_persistent_full.ncu-rep --page raw --csv > $O/p_persistent_full_raw.csv 2>/dev/null
# (c) batch-8 decode: launch list of one graph-free step chain + full capture of the gate|up GEMM
timeout 300 python tools/survey_perf.py decode8 > $O/p_survey8.log 2>&1
FL_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file $O/p_launches_b8.csv \
    python bench.py --workload mistral7b_b8 --steps 4 --warmup 3 --no-cpu > $O/p_ncu_b8.log 2>&1
ls -la $O | tail -30
