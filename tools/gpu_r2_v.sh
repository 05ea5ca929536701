#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 128 --warmup 8 --no-cpu --no-extras"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/v_$name.json 2> $O/v_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/v_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"],1), round(d["ms_per_step"],4), round(d["roofline"]["frac"],4))
except Exception as e: print("$name", "ERR", e)
PY
}
run la0 FL_PK_LOOKAHEAD=0
run gap4 FL_PK_LOOKAHEAD=4 FL_PK_FLAGS=16
run gap8 FL_PK_LOOKAHEAD=8 FL_PK_FLAGS=16
run gap12 FL_PK_LOOKAHEAD=12 FL_PK_FLAGS=16
run gap2 FL_PK_LOOKAHEAD=2 FL_PK_FLAGS=16
ph() { name=$1; shift; echo "== $name"; env "$@" FL_PK_DEBUG=1 timeout 200 python tools/pk_phase_times.py 2>&1 | grep "FL_PK_DEBUG" | tail -6; }
(ph gap8 FL_PK_LOOKAHEAD=8 FL_PK_FLAGS=16) > $O/v_phases.log 2>&1
