set -x
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python tools/run_case.py 100 27 ldlt d 2>&1 | grep -E "factorize|solve|backward" | tail -4
timeout 300 python tools/run_case.py 64 7 llt d 2>&1 | grep -E "factorize|solve|backward" | tail -4
timeout 300 python tools/run_case.py 40 7 ldlh z 2>&1 | grep -E "factorize|solve|backward" | tail -3
