#!/bin/bash
# 1-GPU: fused GEMM tails (correctness + A/B), BERT under PDL, default bench
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rA --timeout 900 > $O/f_pytest.log 2>&1; echo "pytest rc=$?" >> $O/f_pytest.log
timeout 300 python tools/survey_perf.py decode8 decode64 > $O/f_survey.log 2>&1
FL_NO_FUSE=1 timeout 300 python tools/survey_perf.py decode8 decode64 > $O/f_survey_nofuse.log 2>&1
timeout 400 python bench.py --workload mixtral8x7b_b32 --steps 32 --warmup 4 --no-cpu > $O/f_mixtral.json 2> $O/f_mixtral.err
FL_NO_FUSE=1 timeout 400 python bench.py --workload mixtral8x7b_b32 --steps 32 --warmup 4 --no-cpu > $O/f_mixtral_nofuse.json 2> $O/f_mixtral_nofuse.err
timeout 300 python bench.py --workload minilm_256x128 --steps 30 --warmup 5 --no-cpu > $O/f_minilm.json 2> $O/f_minilm.err
FL_NO_PDL=1 timeout 300 python bench.py --workload minilm_256x128 --steps 30 --warmup 5 --no-cpu > $O/f_minilm_nopdl.json 2> $O/f_minilm_nopdl.err
FL_BERT_NO_LNFUSE=1 timeout 300 python bench.py --workload minilm_256x128 --steps 30 --warmup 5 --no-cpu > $O/f_minilm_nolnfuse.json 2> $O/f_minilm_nolnfuse.err
FL_BERT_NO_BRES=1 timeout 300 python bench.py --workload minilm_256x128 --steps 30 --warmup 5 --no-cpu > $O/f_minilm_nobres.json 2> $O/f_minilm_nobres.err
timeout 300 python tools/bert_perf.py > $O/f_bert_perf.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/f_bench_default.json 2> $O/f_bench_default.err; echo "rc=$?" >> $O/f_bench_default.err
