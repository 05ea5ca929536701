#!/bin/bash
# round 2, call x: attention phase of the persistent kernel as its own (not inlined) function with 8 tokens per lane group in flight
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_fulldepth_gpu.py tests/test_parity_gpu.py -m gpu -q -x --timeout 300 > $O/x_pytest.log 2>&1; echo "pytest rc=$?" >> $O/x_pytest.log
tail -3 $O/x_pytest.log
B="python bench.py --steps 128 --warmup 8 --no-cpu --no-extras"
run() { name=$1; shift; env "$@" timeout 200 $B > $O/x_$name.json 2> $O/x_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/x_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"],1), round(d["ms_per_step"],4), round(d["roofline"]["frac"],4), round(d["e2e"]["value"],1))
except Exception as e: print("$name", "ERR", e)
PY
}
run a FL_X=0
run b FL_X=0
(FL_PK_DEBUG=1 timeout 200 python tools/pk_phase_times.py 2>&1 | grep "FL_PK_DEBUG" | grep -v producer | tail -4) > $O/x_phases.log 2>&1
cat $O/x_phases.log
