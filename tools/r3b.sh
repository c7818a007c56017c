#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
T=${1:-r3b}
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_parity_scale.py -m gpu -q -x > gpurun_out/${T}_pytest_gpu_1gpu.log 2>&1; echo rc=$?
tail -4 gpurun_out/${T}_pytest_gpu_1gpu.log
run() { timeout 300 python tools/run_case.py "$@" 2>&1 | grep -E 'factorize|backward|rror' | tail -3; }
for P in 2 0; do
echo "== C2 pdl=$P"; PB200_PDL=$P run 64 7 llt d --reps=4
echo "== C3 pdl=$P"; PB200_PDL=$P run 100 27 ldlt d --reps=2
echo "== c4s pdl=$P"; PB200_PDL=$P run 64 cd lu z --reps=2
done
