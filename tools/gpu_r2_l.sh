#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python tools/survey_perf.py qwen_prefill > $O/l_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_prefill_tc -s 30 -c 1 -o $O/l_attn_tc_full -f python tools/survey_perf.py qwen_prefill > $O/l_ncu.log 2>&1
ncu -i $O/l_attn_tc_full.ncu-rep --page raw --csv > $O/l_attn_tc_full_raw.csv 2>/dev/null
ncu -i $O/l_attn_tc_full.ncu-rep --page source --csv > $O/l_attn_tc_full_source.csv 2>/dev/null
ls -la $O | grep l_
