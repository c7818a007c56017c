#!/bin/bash
# round 2 GPU call: full GPU suite, solve timings, default bench line, ncu evidence with the third-generation up_down sweeps
set -x
export PYTHONUNBUFFERED=1
T=${1:-r2w}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu_1gpu.log 2>&1; echo rc=$?
tail -6 gpurun_out/${T}_pytest_gpu_1gpu.log
run() { timeout 300 python tools/run_case.py "$@" 2>&1 | grep -E 'solve|backward|rror|dag' | tail -3; }
echo "== C2"; run 64 7 llt d --reps=1
echo "== C3"; run 100 27 ldlt d --reps=1
echo "== c4s"; run 64 cd lu z --reps=1
timeout 900 python bench.py > gpurun_out/${T}_bench_c2_n1.json 2> gpurun_out/${T}_bench_c2_n1.err; echo rc=$?
tail -c 1500 gpurun_out/${T}_bench_c2_n1.json
PROFILE_C3=0 bash tools/gpu_profile.sh r02b
