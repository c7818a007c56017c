set -x
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload c5s --steps 3 --no-cpu-baseline > gpurun_out/bench_c5s.json 2> gpurun_out/bench_c5s.log; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_c5s.json')); print('c5s fact_ms', d['fact_ms'], 'solve/rhs', d['solve_ms_per_rhs'], 'launches', d['gpu_launches'], 'e2e numfact', d['e2e']['numfact_call_ms'], 'solve call', d['e2e']['solve_call_ms'], d['backward_error'])"
timeout 300 python tools/run_case.py 64 7 llt d 2>&1 | grep -E "factorize|solve|backward" | tail -3
timeout 300 python tools/run_case.py 48 7 llt s 4 2>&1 | grep -E "factorize|solve|backward" | tail -4
