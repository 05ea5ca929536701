#!/bin/bash
# round 2, final N-GPU check of the default bench line (TP-N + EP-N parity legs) -- usage: gpu_r2_n8.sh <N>
cd "$(dirname "$0")/.." || exit 1
N=${1:-8}
O=gpurun_out; mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 64 --warmup 8 > $O/n${N}_bench_default.json 2> $O/n${N}_bench_default.err; echo "rc=$?" >> $O/n${N}_bench_default.err
tail -1 $O/n${N}_bench_default.err
python - <<PY
import json
d=json.loads(open("$O/n${N}_bench_default.json").read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["e2e"]["value"],1), json.dumps(d.get("parity"))); print([(x["batch"], round(x["value"])) for x in d["batch_sweep"]]); print(round(d["moe"]["value"]), json.dumps(d["moe"]["parity"]))
PY
