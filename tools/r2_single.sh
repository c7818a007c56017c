set -x
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_pytest_gpu.log
tail -8 gpurun_out/r2l_pytest_gpu.log
run() { timeout 300 python tools/run_case.py "$@" 2>&1 | grep -E "factorize|solve|backward|rror" | tail -5; }
echo "== c2s mma"; run 64 7 llt s --reps=3
echo "== c2s simt"; PB200_NO_MMA_SINGLE=1 run 64 7 llt s --reps=3
echo "== c lu 40 mma"; run 40 cd lu c --reps=3
echo "== c lu 40 simt"; PB200_NO_MMA_SINGLE=1 run 40 cd lu c --reps=3
echo "== c3s mma"; run 100 27 ldlt s --reps=2
bash tools/r2_bench.sh c2s 1 --workload c2s
