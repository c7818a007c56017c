#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r3f_pytest_gpu_1gpu.log 2>&1; echo rc=$?
tail -4 gpurun_out/r3f_pytest_gpu_1gpu.log
run() { timeout 300 python tools/run_case.py "$@" 2>&1 | grep -E 'factorize|backward' | tail -3; }
for P in 1 0; do
echo "== C2 cmp=$P"; PB200_DIAG_CMP=$P run 64 7 llt d --reps=3
echo "== C3 cmp=$P"; PB200_DIAG_CMP=$P run 100 27 ldlt d --reps=2
echo "== c2s cmp=$P"; PB200_DIAG_CMP=$P run 64 7 llt s --reps=2
done
