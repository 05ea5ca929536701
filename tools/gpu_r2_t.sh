#!/bin/bash
# round 2, call t: L2 prefetch thread in the persistent kernel (HBM keeps streaming through the grid barriers / attention phase)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
echo skip-pytest > $O/t_pytest.log
tail -3 $O/t_pytest.log
B="python bench.py --steps 128 --warmup 8 --no-cpu --no-extras"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/t_$name.json 2> $O/t_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/t_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"],1), round(d["ms_per_step"],4), round(d["roofline"]["frac"],4))
except Exception as e: print("$name", "ERR", e)
PY
}
run la0 FL_PK_LOOKAHEAD=0
run la4 FL_PK_LOOKAHEAD=4
run la8 FL_PK_LOOKAHEAD=8
run la2 FL_PK_LOOKAHEAD=2
run la16 FL_PK_LOOKAHEAD=16
run la6 FL_PK_LOOKAHEAD=6

run la8_static FL_PK_LOOKAHEAD=8 FL_PK_STATIC=32

ph() { name=$1; shift; echo "== $name"; env "$@" FL_PK_DEBUG=1 timeout 200 python tools/pk_phase_times.py 2>&1 | grep "FL_PK_DEBUG" | tail -6; }
(ph la8 FL_PK_LOOKAHEAD=8; ph la4 FL_PK_LOOKAHEAD=4) > $O/t_phases.log 2>&1
