#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python tools/diag_batch3.py > $O/b_diag.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rA --timeout 900 > $O/b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/b_pytest.log
for d in 0 1 2 3; do
  FL_GEMM_DBG=$d timeout 300 python tools/survey_perf.py decode8 decode64 > $O/b_survey_dbg$d.log 2>&1
done
