set -x
export PYTHONUNBUFFERED=1
PB200_PROFILE=1 PB200_PROFILE_VERBOSE=1 timeout 300 python tools/run_case.py 64 7 llt d > gpurun_out/levels_c2.txt 2>&1
PB200_PROFILE=1 PB200_PROFILE_VERBOSE=1 timeout 400 python tools/run_case.py 100 27 ldlt d > gpurun_out/levels_c3.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_c2.csv python tools/profile_step.py c2 > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_gemm_scatter -s 24 -c 8 -o /tmp/prof_gs python tools/profile_step.py c2 > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_gs.ncu-rep --page raw --csv > gpurun_out/prof_gemm_scatter_c2_raw.csv 2>/dev/null
ncu -i /tmp/prof_gs.ncu-rep --page details --csv > gpurun_out/prof_gemm_scatter_c2_details.csv 2>/dev/null
ncu -i /tmp/prof_gs.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_gemm_scatter_c2_source.csv 2>/dev/null
ls -la /tmp/prof_gs.ncu-rep
SZ=$(stat -c %s /tmp/prof_gs.ncu-rep); if [ "$SZ" -lt 30000000 ]; then cp /tmp/prof_gs.ncu-rep gpurun_out/prof_gemm_scatter_c2.ncu-rep; fi
for bs in "120 240" "240 480"; do set -- $bs
  timeout 300 python bench.py --workload c2 --steps 3 --no-cpu-baseline --iparm IPARM_MIN_BLOCKSIZE=$1 --iparm IPARM_MAX_BLOCKSIZE=$2 > gpurun_out/bench_c2_bs$2.json 2> gpurun_out/bench_c2_bs$2.log
  timeout 400 python bench.py --workload c3 --steps 3 --no-cpu-baseline --iparm IPARM_MIN_BLOCKSIZE=$1 --iparm IPARM_MAX_BLOCKSIZE=$2 > gpurun_out/bench_c3_bs$2.json 2> gpurun_out/bench_c3_bs$2.log
done
cat gpurun_out/bench_c*_bs*.json
du -sh gpurun_out; ls -la gpurun_out
