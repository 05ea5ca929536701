#!/bin/bash
# round 2, final 2-GPU check: TP / EP tests + the default bench line at N = 2 (with its parity legs)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_tp.py -m gpu -q -rA --timeout 600 > $O/tp2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/tp2_pytest.log
tail -6 $O/tp2_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 > $O/tp2_bench_default.json 2> $O/tp2_bench_default.err; echo "rc=$?" >> $O/tp2_bench_default.err
tail -2 $O/tp2_bench_default.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --workload qwen25_7b_prefill4k --steps 4 --warmup 3 --no-cpu > $O/tp2_qwen_prefill.json 2> $O/tp2_qwen_prefill.err; echo "rc=$?" >> $O/tp2_qwen_prefill.err
python - <<PY
import json
for f in ["tp2_bench_default","tp2_qwen_prefill"]:
    try:
        d=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],4), d.get("parity"))
    except Exception as e: print(f, "ERR", e)
PY
