#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
T=r3g
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu_1gpu.log 2>&1; echo rc=$?
tail -4 gpurun_out/${T}_pytest_gpu_1gpu.log
run() { timeout 300 python tools/run_case.py "$@" 2>&1 | grep -E 'solve|backward|rror' | tail -3; }
echo "== C2"; run 64 7 llt d --reps=1
echo "== C3"; run 100 27 ldlt d --reps=1
echo "== c4s"; run 64 cd lu z --reps=1
echo "== c2s"; run 64 7 llt s --reps=1
echo "== C2 trace"; PB200_DAG_TRACE=gpurun_out/${T}_trace_c2.bin run 64 7 llt d --reps=1
python tools/dag_trace.py gpurun_out/${T}_trace_c2.bin
gzip -f gpurun_out/${T}_trace_c2.bin
