#!/bin/bash
# round 2: eight-rows-per-thread q|k|v epilogue in prefill -- full GPU suite + A/B on the Qwen2.5-7B 4k prefill
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rA --timeout 900 > $O/qr_pytest.log 2>&1; echo "pytest rc=$?" >> $O/qr_pytest.log
tail -2 $O/qr_pytest.log
FL_QKV_ROWS=0 timeout 600 python bench.py --workload qwen25_7b_prefill4k --steps 8 --warmup 3 --no-cpu > $O/qr_prefill_old.json 2> $O/qr_prefill_old.err
timeout 600 python bench.py --workload qwen25_7b_prefill4k --steps 8 --warmup 3 --no-cpu > $O/qr_prefill_new.json 2> $O/qr_prefill_new.err
python - <<PY
import json
for f in ["qr_prefill_old","qr_prefill_new"]:
    try:
        d=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1]); k=[x for x in d["roofline"]["per_kernel"] if x["kernel"] in ("dense_qkv_rope_append","attn_prefill_tc","dense_resid_rmsnorm")]
        print(f, round(d["value"],1), round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], [(x["kernel"], round(x["ms"],3)) for x in k])
    except Exception as e: print(f, "ERR", e)
PY
