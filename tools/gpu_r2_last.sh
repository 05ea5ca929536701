#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python __graft_entry__.py smoke > $O/last_smoke.log 2>&1; echo "smoke rc=$?" >> $O/last_smoke.log
timeout 900 python bench.py > $O/last_bench_default.json 2> $O/last_bench_default.err; echo "rc=$?" >> $O/last_bench_default.err
tail -2 $O/last_smoke.log; tail -1 $O/last_bench_default.err
python - <<PY
import json
d=json.loads(open("$O/last_bench_default.json").read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["roofline"]["frac"],4), round(d["e2e"]["value"],1), d["parity"]["greedy32"], [(x["batch"], round(x["value"])) for x in d["batch_sweep"]], round(d["moe"]["value"]), round(d["secondary"]["value"]))
PY
