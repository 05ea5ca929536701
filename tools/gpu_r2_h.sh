#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rA --timeout 900 > $O/h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/h_pytest.log
timeout 300 python bench.py --workload minilm_256x128 --steps 200 --warmup 20 --no-cpu > $O/h_minilm.json 2> $O/h_minilm.err
timeout 300 python bench.py --workload minilm_256x128 --steps 200 --warmup 20 --no-cpu > $O/h_minilm2.json 2> $O/h_minilm2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file $O/h_launches_minilm.csv \
    python bench.py --workload minilm_256x128 --steps 4 --warmup 3 --no-cpu > $O/h_ncu_minilm.log 2>&1
timeout 300 python bench.py --steps 64 --warmup 8 --no-cpu --no-extras > $O/h_b1.json 2> $O/h_b1.err
timeout 300 python bench.py --steps 64 --warmup 8 --no-cpu --no-extras > $O/h_b1_2.json 2> $O/h_b1_2.err
timeout 300 python tools/survey_perf.py decode8 decode64 > $O/h_survey.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/h_bench_default.json 2> $O/h_bench_default.err; echo "rc=$?" >> $O/h_bench_default.err
timeout 300 python __graft_entry__.py smoke > $O/h_smoke.log 2>&1; echo "smoke rc=$?" >> $O/h_smoke.log
