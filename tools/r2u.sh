#!/bin/bash
# round 2 GPU call: k_dag3 — column-split chain tickets, A/B against the first generation
set -x
export PYTHONUNBUFFERED=1
T=${1:-r2v}
PB200_DAG3=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_parity_scale.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo rc=$?
tail -3 gpurun_out/${T}_pytest.log
run() { timeout 300 python tools/run_case.py "$@" 2>&1 | grep -E 'solve|backward|rror|dag' | tail -3; }
for CS in 0 600 2000; do
echo "== C2 v3 csplit $CS"; PB200_DAG3=1 PB200_DAG3_CSPLIT=$CS run 64 7 llt d --reps=1
echo "== C3 v3 csplit $CS"; PB200_DAG3=1 PB200_DAG3_CSPLIT=$CS run 100 27 ldlt d --reps=1
done
echo "== C2 v1"; PB200_DAG3=0 run 64 7 llt d --reps=1
echo "== C2 trace"; PB200_DAG3=1 PB200_DAG_TRACE=gpurun_out/${T}_trace_c2.bin run 64 7 llt d --reps=1
python tools/dag_trace.py gpurun_out/${T}_trace_c2.bin
gzip -f gpurun_out/${T}_trace_c2.bin
echo "== c4s v3"; PB200_DAG3=1 run 64 cd lu z --reps=1
echo "== c4s v1"; PB200_DAG3=0 run 64 cd lu z --reps=1
