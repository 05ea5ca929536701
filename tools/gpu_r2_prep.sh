#!/bin/bash
# round 2: RMSNorm prologue with four split-K slice loads in flight -- full GPU suite on the build of record
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -rA --timeout 600 > $O/prep_pytest.log 2>&1; echo "pytest rc=$?" >> $O/prep_pytest.log
tail -2 $O/prep_pytest.log
