#!/bin/bash
# round 2, call n: ring-stage lending in the persistent kernel (6 stages instead of 4 on Mistral-7B) + whole-item attention loads
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_fulldepth_gpu.py tests/test_parity_gpu.py -m gpu -q -x --timeout 600 > $O/n_pytest.log 2>&1; echo "pytest rc=$?" >> $O/n_pytest.log
tail -3 $O/n_pytest.log
B="python bench.py --steps 128 --warmup 8 --no-cpu --no-extras"
timeout 300 $B > $O/n_lend_b8.json 2> $O/n_lend_b8.err
FL_PK_NOLEND=1 timeout 300 $B > $O/n_nolend_b8.json 2> $O/n_nolend_b8.err
FL_PK_DEBUG=1 timeout 200 python tools/pk_phase_times.py > $O/n_phase_lend_b8.log 2>&1
cp fastllm_b200/libfastllm_b200.so /tmp/lib_keep.so
cp tools/_build/lib_batch4.so fastllm_b200/libfastllm_b200.so
timeout 300 $B > $O/n_lend_b4.json 2> $O/n_lend_b4.err
FL_PK_DEBUG=1 timeout 200 python tools/pk_phase_times.py > $O/n_phase_lend_b4.log 2>&1
cp /tmp/lib_keep.so fastllm_b200/libfastllm_b200.so
for f in n_lend_b8 n_nolend_b8 n_lend_b4; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"])
except Exception as e: print("$f", "ERR", e)
PY
done
grep -h "layer total" $O/n_phase_lend_b8.log | tail -2
grep -h "layer total" $O/n_phase_lend_b4.log | tail -2
