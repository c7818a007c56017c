#!/bin/bash
# round 2 GPU call: full GPU suite, default bench line, ncu evidence with the third-generation up_down sweeps
set -x
export PYTHONUNBUFFERED=1
T=${1:-r2w}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu_1gpu.log 2>&1; echo rc=$?
tail -6 gpurun_out/${T}_pytest_gpu_1gpu.log
timeout 900 python bench.py > gpurun_out/${T}_bench_c2_n1.json 2> gpurun_out/${T}_bench_c2_n1.err; echo rc=$?
tail -c 3000 gpurun_out/${T}_bench_c2_n1.json
PROFILE_C3=0 bash tools/gpu_profile.sh r02b
