set -x
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for v in "" vE vF; do
  echo "== variant '$v'"
  if [ -n "$v" ]; then export PB200_LIB=$PWD/pastix_b200/lib/libpastix_b200_$v.so; else unset PB200_LIB; fi
  timeout 300 python tools/run_case.py 64 7 llt d 2>&1 | grep -E "factorize|backward" | tail -2
  timeout 300 python tools/run_case.py 100 27 ldlt d 2>&1 | grep -E "factorize|backward" | tail -2
done > gpurun_out/variants.log 2>&1
unset PB200_LIB
cat gpurun_out/variants.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.log; echo "rc=$?"; cat gpurun_out/bench_c2.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c2_ref.json 2> gpurun_out/bench_c2_ref.log; echo "rc=$?"; cat gpurun_out/bench_c2_ref.json
