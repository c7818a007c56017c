# GPU test suite with its log kept (profiles/r02/): usage tools/r2_tests.sh <tag>
export PYTHONUNBUFFERED=1
TAG=${1:-x}
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r2_pytest_gpu_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu_$TAG.log
tail -40 gpurun_out/r2_pytest_gpu_$TAG.log
