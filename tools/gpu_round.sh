set -x
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.log; echo "bench rc=$?"
cat gpurun_out/bench_c2.json
timeout 600 python bench.py --workload c3 --steps 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.log; echo "bench c3 rc=$?"
cat gpurun_out/bench_c3.json
timeout 300 python tools/fp64_probe.py > gpurun_out/fp64_probe.txt 2>&1
cat gpurun_out/fp64_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_c2.csv python tools/profile_step.py c2 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_gemm_scatter -o gpurun_out/prof_gemm_scatter_c2 python tools/profile_step.py c2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
