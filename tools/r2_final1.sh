export PYTHONUNBUFFERED=1
bash tools/r2_tests.sh final_n1
bash tools/r2_bench.sh final_n1 1
bash tools/r2_bench.sh c3_tuned 1 --workload c3 --tuned
bash tools/r2_bench.sh c2_tuned 1 --workload c2 --tuned
timeout 600 python bench.py --impl reference --workload c2 --tuned --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_c2_tuned.json 2>/dev/null; cat gpurun_out/r2_bench_ref_c2_tuned.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_c2.json 2>/dev/null; cat gpurun_out/r2_bench_ref_c2.json | cut -c1-300
