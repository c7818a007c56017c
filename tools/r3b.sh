#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
T=r3o
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu_1gpu.log 2>&1; echo rc=$?
tail -3 gpurun_out/${T}_pytest_gpu_1gpu.log
run() { timeout 300 python tools/run_case.py "$@" 2>&1 | grep -E 'factorize|solve|backward' | tail -3; }
echo "== C2 pdl=0"; PB200_PDL=0 run 64 7 llt d --reps=3
echo "== C2"; run 64 7 llt d --reps=3
echo "== c4s"; run 64 cd lu z --reps=2
timeout 900 python bench.py > gpurun_out/${T}_bench_c2_n1.json 2> gpurun_out/${T}_bench_c2_n1.err; echo rc=$?
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_ref_c2.json 2> gpurun_out/${T}_bench_ref_c2.err; echo rc=$?
tail -c 300 gpurun_out/${T}_bench_ref_c2.json
