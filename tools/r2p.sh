#!/bin/bash
# round 2 GPU call: second-generation up_down sweeps (k_dag2) — parity, A/B against the first generation, traces
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_parity_scale.py tests/test_dropin_gpu.py -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo rc=$?
tail -8 gpurun_out/r2p_pytest.log
run() { timeout 300 python tools/run_case.py "$@" 2>&1 | grep -E 'factorize|solve|backward|rror|dag' | tail -5; }
echo "== C2 v2"; PB200_DAG_VERBOSE=1 run 64 7 llt d --reps=1
echo "== C2 v1"; PB200_DAG_V1=1 run 64 7 llt d --reps=1
echo "== C2 v2 trace"; PB200_DAG_TRACE=gpurun_out/r2p_trace_c2.bin run 64 7 llt d --reps=1
python tools/dag_trace.py gpurun_out/r2p_trace_c2.bin
echo "== C2 v2 64 rhs"; run 64 7 llt d 64 --reps=1
echo "== C2 v1 64 rhs"; PB200_DAG_V1=1 run 64 7 llt d 64 --reps=1
echo "== C3 v2"; run 100 27 ldlt d --reps=1
echo "== C3 v1"; PB200_DAG_V1=1 run 100 27 ldlt d --reps=1
echo "== C3 v2 trace"; PB200_DAG_TRACE=gpurun_out/r2p_trace_c3.bin run 100 27 ldlt d --reps=1
python tools/dag_trace.py gpurun_out/r2p_trace_c3.bin
gzip -f gpurun_out/r2p_trace_c2.bin; rm -f gpurun_out/r2p_trace_c3.bin
echo "== c4s v2"; run 64 cd lu z --reps=1
echo "== c4s v1"; PB200_DAG_V1=1 run 64 cd lu z --reps=1
