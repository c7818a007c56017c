#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
T=${1:-r2z}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu_1gpu.log 2>&1; echo rc=$?
tail -6 gpurun_out/${T}_pytest_gpu_1gpu.log
PB200_SHIM_TIMING=1 timeout 900 python bench.py > gpurun_out/${T}_bench_c2_n1.json 2> gpurun_out/${T}_bench_c2_n1.err; echo rc=$?
grep "pb200 shim" gpurun_out/${T}_bench_c2_n1.err | tail -4
python - <<'P'
import json,sys
d=json.loads(open(sys.argv[1] if len(sys.argv)>1 else "gpurun_out/%s_bench_c2_n1.json" % "T").read().strip().splitlines()[-1]) if False else None
P
tail -c 600 gpurun_out/${T}_bench_c2_n1.json
