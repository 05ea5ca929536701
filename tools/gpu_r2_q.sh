#!/bin/bash
# round 2, call q: persistent kernel with the shortened consumer chain (no integer division per chunk, shared-space addresses hoisted,
# all fragment loads of a chunk ahead of four MMA chains)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_fulldepth_gpu.py tests/test_parity_gpu.py -m gpu -q -x --timeout 600 > $O/q_pytest.log 2>&1; echo "pytest rc=$?" >> $O/q_pytest.log
tail -3 $O/q_pytest.log
B="python bench.py --steps 128 --warmup 8 --no-cpu --no-extras"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/q_$name.json 2> $O/q_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/q_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"],1), round(d["ms_per_step"],4), round(d["roofline"]["frac"],4))
except Exception as e: print("$name", "ERR", e)
PY
}
run ns4 FL_X=0
run ns3 FL_PK_MAXSTAGES=3
run ns2 FL_PK_MAXSTAGES=2
run noload FL_PK_FLAGS=8
run skeleton FL_PK_FLAGS=14
ph() { name=$1; shift; echo "== $name"; env "$@" FL_PK_DEBUG=1 timeout 200 python tools/pk_phase_times.py 2>&1 | grep "FL_PK_DEBUG" | tail -2; }
(ph ns4 FL_X=1; ph noload FL_PK_FLAGS=8; ph skeleton FL_PK_FLAGS=14) > $O/q_phases.log 2>&1
