#!/bin/bash
# Dev tool: the single-GPU evidence pass a round starts (and ends) with, as ONE gpurun call:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round.sh rNN'
# 1. pytest -m gpu (parity through the C ABI), smoke()   2. the bench lines of the single-GPU BASELINE configs
# 3. A/B of the experimental knobs (parity suite under the knob first, then the bench lines it is meant to move)
# 4. ncu launch list of the default bench + one --set full capture of the persistent decode kernel (only after 1-2 exited 0).
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/.
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
set -x
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/${TAG}_smoke.log
bench() { # name, args...
  name=$1; shift
  python bench.py "$@" > $OUT/${TAG}_bench_${name}.json 2> $OUT/${TAG}_bench_${name}.err; echo "bench $name rc=$?"
}
bench default
bench b8 --workload mistral7b_b8 --steps 64
bench b64 --workload mistral7b_b64 --steps 32
bench tinyllama --workload tinyllama_b1
bench qwen_prefill --workload qwen25_7b_prefill4k --steps 4
bench minilm --workload minilm_256x128 --steps 20
# experimental: weight tiles of the swap-AB decode GEMMs requested ahead of the dependency wait (csrc/gemm_tc.cuh producer)
FL_GEMM_WPREFETCH=1 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu_wprefetch.log 2>&1; echo "pytest(wprefetch) rc=$?" | tee -a $OUT/${TAG}_pytest_gpu_wprefetch.log
if tail -1 $OUT/${TAG}_pytest_gpu_wprefetch.log | grep -q "rc=0"; then
  FL_GEMM_WPREFETCH=1 bench b8_wprefetch --workload mistral7b_b8 --steps 64
  FL_GEMM_WPREFETCH=1 bench b64_wprefetch --workload mistral7b_b64 --steps 32
fi
# profiler passes last: numbers printed under ncu are never bench values
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_default.csv \
    python bench.py --steps 2 --warmup 1 > $OUT/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_persistent -c 1 -o $OUT/${TAG}_persistent_full -f \
    python bench.py --steps 2 --warmup 1 > $OUT/${TAG}_ncu_full.log 2>&1
ncu -i $OUT/${TAG}_persistent_full.ncu-rep --page raw --csv > $OUT/${TAG}_persistent_full_raw.csv 2>/dev/null
ls -la $OUT | tail -30
