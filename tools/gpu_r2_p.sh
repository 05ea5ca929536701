#!/bin/bash
# round 2, call p: ring-depth sensitivity of the persistent kernel (2 / 3 / 4 stages, 6 with lending) + phase times under the experiments
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 128 --warmup 8 --no-cpu --no-extras"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/p_$name.json 2> $O/p_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/p_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"],1), round(d["ms_per_step"],4), round(d["roofline"]["frac"],4))
except Exception as e: print("$name", "ERR", e)
PY
}
run ns2 FL_PK_MAXSTAGES=2
run ns3 FL_PK_MAXSTAGES=3
run ns4 FL_PK_MAXSTAGES=4
run ns6lend FL_PK_LEND=1
ph() { name=$1; shift; echo "== $name"; env "$@" FL_PK_DEBUG=1 timeout 200 python tools/pk_phase_times.py 2>&1 | grep "FL_PK_DEBUG" | tail -2; }
ph ns2 FL_PK_MAXSTAGES=2
ph ns4 FL_X=1
ph ns6lend FL_PK_LEND=1
ph ns4_nomath FL_PK_FLAGS=2
ph ns4_noload FL_PK_FLAGS=8
ph ns4_skeleton FL_PK_FLAGS=14
