#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 > $O/tp2_bench_default.json 2> $O/tp2_bench_default.err; echo "rc=$?" >> $O/tp2_bench_default.err
tail -1 $O/tp2_bench_default.err
python - <<PY
import json
d=json.loads(open("$O/tp2_bench_default.json").read().strip().splitlines()[-1]); print(round(d["value"],1), d.get("parity",{}).get("greedy32")); print(json.dumps(d["moe"]["parity"]))
PY
