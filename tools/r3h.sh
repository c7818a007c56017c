#!/bin/bash
# final 1-GPU verification of round 2: suite, sanitizer on the persistent sweeps, bench line, ncu evidence
set -x
export PYTHONUNBUFFERED=1
T=r3h
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu_1gpu.log 2>&1; echo rc=$?
tail -4 gpurun_out/${T}_pytest_gpu_1gpu.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "second_generation and (lap7_8_llt_d or cd_6_lu_z or lap27_6_ldlt_d)" > gpurun_out/${T}_memcheck.log 2>&1; echo memcheck rc=$?
tail -6 gpurun_out/${T}_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "second_generation and lap7_8_llt_d and 1" > gpurun_out/${T}_racecheck.log 2>&1; echo racecheck rc=$?
tail -6 gpurun_out/${T}_racecheck.log
timeout 900 python bench.py > gpurun_out/${T}_bench_c2_n1.json 2> gpurun_out/${T}_bench_c2_n1.err; echo rc=$?
PROFILE_C3=0 bash tools/gpu_profile.sh r02c
