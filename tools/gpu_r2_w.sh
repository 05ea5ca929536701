#!/bin/bash
# round 2, call w: decode_persistent_ll_kernel (no grid barriers inside a layer) -- correctness, then A/B against the barrier kernel
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_fulldepth_gpu.py tests/test_parity_gpu.py -m gpu -q -x --timeout 300 > $O/w_pytest.log 2>&1; echo "pytest rc=$?" >> $O/w_pytest.log
tail -5 $O/w_pytest.log
B="python bench.py --steps 128 --warmup 8 --no-cpu --no-extras"
run() { name=$1; shift; env "$@" timeout 200 $B > $O/w_$name.json 2> $O/w_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/w_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"],1), round(d["ms_per_step"],4), round(d["roofline"]["frac"],4), round(d["e2e"]["value"],1))
except Exception as e: print("$name", "ERR", e)
PY
}
run ll FL_X=0
run noll FL_PK_NOLL=1
(FL_PK_DEBUG=1 timeout 200 python tools/pk_phase_times.py 2>&1 | grep "FL_PK_DEBUG" | grep -v producer | tail -4) > $O/w_phases.log 2>&1
cat $O/w_phases.log
