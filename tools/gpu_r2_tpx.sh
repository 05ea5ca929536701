#!/bin/bash
# round 2: fused TP exchange (TPX) -- correctness at N GPUs, A/B against exchange + barrier, and the 1-GPU headline with the new build
cd "$(dirname "$0")/.." || exit 1
N=${1:-2}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_tp.py -m gpu -q -rA --timeout 600 > $O/tpx_pytest_n$N.log 2>&1; echo "pytest rc=$?" >> $O/tpx_pytest_n$N.log
tail -8 $O/tpx_pytest_n$N.log | cut -c1-200
run() { name=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 128 --warmup 8 --no-extras --no-cpu > $O/tpx_${name}_n$N.json 2> $O/tpx_${name}_n$N.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/tpx_${name}_n$N.json").read().strip().splitlines()[-1]); print("$name", round(d["value"],1), round(d["ms_per_step"],4), d.get("parity",{}).get("greedy32"), d.get("parity",{}).get("max_abs"))
except Exception as e: print("$name", "ERR", e)
PY
}
run fused FL_X=0
run unfused FL_PK_TP_FUSED=0
timeout 300 python bench.py --steps 128 --warmup 8 --no-extras --no-cpu > $O/tpx_single.json 2> $O/tpx_single.err
python - <<PY
import json
d=json.loads(open("$O/tpx_single.json").read().strip().splitlines()[-1]); print("single", round(d["value"],1), round(d["ms_per_step"],4), round(d["roofline"]["frac"],4))
PY
