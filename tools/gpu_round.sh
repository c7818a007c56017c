set -x
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python tools/run_case.py 64 7 llt d 2>&1 | grep -E "factorize|solve|backward" | tail -4
timeout 300 python tools/run_case.py 100 27 ldlt d 2>&1 | grep -E "factorize|solve|backward" | tail -4
timeout 300 python tools/run_case.py 48 cd lu z 2>&1 | grep -E "factorize|solve|backward|analysis" | tail -5
timeout 300 python tools/run_case.py 64 7 llt d 8 2>&1 | grep -E "solve" | tail -2
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_gemm_scatter -s 300 -c 4 -o /tmp/prof_gs3 python tools/profile_step.py c3 > gpurun_out/ncu_full_c3.log 2>&1
ncu -i /tmp/prof_gs3.ncu-rep --page raw --csv > gpurun_out/prof_gemm_scatter_c3_raw.csv 2>/dev/null
ncu -i /tmp/prof_gs3.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_gemm_scatter_c3_source.csv 2>/dev/null
ls -la /tmp/prof_gs3.ncu-rep
