#!/bin/bash
# round 2, call o: what bounds the persistent kernel?  timing experiments (FL_PK_FLAGS bits 1-3: garbage results) + lending A/B
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 128 --warmup 8 --no-cpu --no-extras"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/o_$name.json 2> $O/o_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/o_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"],1), round(d["ms_per_step"],4), round(d["roofline"]["frac"],4))
except Exception as e: print("$name", "ERR", e)
PY
}
run lend FL_X=0
run nolend FL_PK_NOLEND=1
run lend_nomath FL_PK_FLAGS=2
run lend_noattn FL_PK_FLAGS=4
run lend_noload FL_PK_FLAGS=8
run lend_nomath_noattn FL_PK_FLAGS=6
run lend_skeleton FL_PK_FLAGS=14
run nolend_nomath FL_PK_FLAGS=2 FL_PK_NOLEND=1
run lend2 FL_X=0
