#!/bin/bash
# round 2, final profiler pass (numbers printed under ncu are never bench values): launch list of the default command + one full
# capture of the stream-K batched-decode attention at batch 64
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/ncu_launches_default.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_launches_default.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_sk_decode -s 40 -c 1 -o $O/ncu_attn_sk_b64 -f \
    python bench.py --workload mistral7b_b64 --steps 2 --warmup 3 --no-cpu --no-extras > $O/ncu_attn_sk_b64.log 2>&1
echo "full capture rc=$?"
ncu -i $O/ncu_attn_sk_b64.ncu-rep --page raw --csv > $O/ncu_attn_sk_b64_raw.csv 2>/dev/null
wc -l $O/ncu_launches_default.csv $O/ncu_attn_sk_b64_raw.csv
