set -x
export PYTHONUNBUFFERED=1
NB64=pastix_b200/lib/libpastix_b200_nb64.so
PB200_PROFILE=1 PB200_PROFILE_VERBOSE=1 timeout 300 python tools/run_case.py 64 7 llt d --reps=2 > gpurun_out/r2b_levels_c2_nb128.txt 2>&1
PB200_LIB=$NB64 PB200_PROFILE=1 PB200_PROFILE_VERBOSE=1 timeout 300 python tools/run_case.py 64 7 llt d --reps=2 > gpurun_out/r2b_levels_c2_nb64.txt 2>&1
PB200_LIB=$NB64 PB200_DIAG_OLD=1 PB200_PROFILE=1 PB200_PROFILE_VERBOSE=1 timeout 300 python tools/run_case.py 64 7 llt d --reps=2 > gpurun_out/r2b_levels_c2_nb64_old.txt 2>&1
PB200_PROFILE=1 PB200_PROFILE_VERBOSE=1 timeout 300 python tools/run_case.py 100 27 ldlt d --reps=2 > gpurun_out/r2b_levels_c3_nb128.txt 2>&1
PB200_LIB=$NB64 PB200_PROFILE=1 PB200_PROFILE_VERBOSE=1 timeout 300 python tools/run_case.py 100 27 ldlt d --reps=2 > gpurun_out/r2b_levels_c3_nb64.txt 2>&1
grep "pb200 profile" gpurun_out/r2b_levels_*.txt
