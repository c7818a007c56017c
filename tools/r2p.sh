#!/bin/bash
# round 2 GPU call: second-generation up_down sweeps (k_dag2) — parity, traces (first generation measured in r2p: C2 2.5, C3 12.2, c4s 4.05 ms)
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T=${1:-r2q}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_parity_scale.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo rc=$?
tail -8 gpurun_out/${T}_pytest.log
run() { timeout 300 python tools/run_case.py "$@" 2>&1 | grep -E 'factorize|solve|backward|rror|dag' | tail -5; }
echo "== C2 v2"; PB200_DAG_VERBOSE=1 run 64 7 llt d --reps=1
echo "== C2 v2 trace"; PB200_DAG_TRACE=gpurun_out/${T}_trace_c2.bin run 64 7 llt d --reps=1
python tools/dag_trace.py gpurun_out/${T}_trace_c2.bin
echo "== C2 v2 64 rhs"; PB200_DAG_V2=1 run 64 7 llt d 64 --reps=1; echo "== C2 64 rhs default"; run 64 7 llt d 64 --reps=1
echo "== C3 v2"; run 100 27 ldlt d --reps=1
echo "== C3 v2 trace"; PB200_DAG_TRACE=gpurun_out/${T}_trace_c3.bin run 100 27 ldlt d --reps=1
python tools/dag_trace.py gpurun_out/${T}_trace_c3.bin
gzip -f gpurun_out/${T}_trace_c2.bin; rm -f gpurun_out/${T}_trace_c3.bin
echo "== c4s v2"; run 64 cd lu z --reps=1
