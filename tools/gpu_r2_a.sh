#!/bin/bash
# round 2, GPU call A: parity suite + smoke + default bench + A/B knobs.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > $O/a_smi.txt 2>&1
free -g > $O/a_host.txt; nproc >> $O/a_host.txt
timeout 1500 python -m pytest tests -m gpu -q -rA --timeout 900 > $O/a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/a_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/a_smoke.log 2>&1; echo "smoke rc=$?" >> $O/a_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/a_bench_default.json 2> $O/a_bench_default.err; echo "rc=$?" >> $O/a_bench_default.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/a_bench_reference.json 2> $O/a_bench_reference.err; echo "rc=$?" >> $O/a_bench_reference.err
for wl in mistral7b_b8 mistral7b_b64; do
  timeout 300 python bench.py --workload $wl --steps 64 --warmup 8 --no-cpu > $O/a_${wl}.json 2> $O/a_${wl}.err
  FL_GEMM_WPREFETCH=1 timeout 300 python bench.py --workload $wl --steps 64 --warmup 8 --no-cpu > $O/a_${wl}_wprefetch.json 2> $O/a_${wl}_wprefetch.err
done
FL_MOE_MASKED=1 timeout 400 python bench.py --workload mixtral8x7b_b32 --steps 32 --warmup 4 --no-cpu > $O/a_mixtral_masked.json 2> $O/a_mixtral_masked.err
timeout 400 python bench.py --workload mixtral8x7b_b32 --steps 32 --warmup 4 --no-cpu > $O/a_mixtral_grouped.json 2> $O/a_mixtral_grouped.err
echo done > $O/a_done.txt
