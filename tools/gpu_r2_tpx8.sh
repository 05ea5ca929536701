#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
N=${1:-8}
O=gpurun_out; mkdir -p $O
run() { name=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 128 --warmup 8 --no-extras --no-cpu > $O/tpx_${name}_n$N.json 2> $O/tpx_${name}_n$N.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/tpx_${name}_n$N.json").read().strip().splitlines()[-1]); print("$name", round(d["value"],1), round(d["ms_per_step"],4), d.get("parity",{}).get("greedy32"), d.get("parity",{}).get("max_abs"))
except Exception as e: print("$name", "ERR", e)
PY
}
run fused FL_X=0
run unfused FL_PK_TP_FUSED=0
run fused2 FL_X=0
