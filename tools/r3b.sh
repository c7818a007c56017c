#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_parity_scale.py tests/test_dropin_gpu.py -m gpu -q -x > gpurun_out/r3m_pytest.log 2>&1; echo rc=$?
tail -3 gpurun_out/r3m_pytest.log
run() { timeout 300 python tools/run_case.py "$@" 2>&1 | grep -E 'solve|backward' | tail -2; }
echo "== C2"; run 64 7 llt d --reps=1
echo "== C3"; run 100 27 ldlt d --reps=1
echo "== c4s"; run 64 cd lu z --reps=1
echo "== c2s"; run 64 7 llt s --reps=1
