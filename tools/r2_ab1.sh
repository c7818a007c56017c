# round 2, batch 1: TMA-staged k_gemm_scatter + blocked diagonal kernel (single round per cblk) parity and A/B
set -x
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_pytest_gpu.log
tail -15 gpurun_out/r2a_pytest_gpu.log
run() { timeout 300 python tools/run_case.py "$@" 2>&1 | grep -E "factorize|solve|backward|rror" | tail -6; }
NB64=pastix_b200/lib/libpastix_b200_nb64.so
NOTMA=pastix_b200/lib/libpastix_b200_notma.so
echo "== default C2"; run 64 7 llt d --reps=4
echo "== default+graph C2"; PB200_GRAPH=1 run 64 7 llt d --reps=4
echo "== nb64+olddiag C2"; PB200_LIB=$NB64 PB200_DIAG_OLD=1 run 64 7 llt d --reps=4
echo "== nb64 newdiag C2"; PB200_LIB=$NB64 run 64 7 llt d --reps=4
echo "== noTMA C2"; PB200_LIB=$NOTMA run 64 7 llt d --reps=4
echo "== default C3"; run 100 27 ldlt d --reps=3
echo "== default+graph C3"; PB200_GRAPH=1 run 100 27 ldlt d --reps=3
echo "== nb64+olddiag C3"; PB200_LIB=$NB64 PB200_DIAG_OLD=1 run 100 27 ldlt d --reps=3
echo "== noTMA C3"; PB200_LIB=$NOTMA run 100 27 ldlt d --reps=3
echo "== default z-LU 40"; run 40 cd lu z --reps=3
echo "== nb64+olddiag z-LU 40"; PB200_LIB=$NB64 PB200_DIAG_OLD=1 run 40 cd lu z --reps=3
echo "== default d-LU 48"; run 48 cd lu d --reps=3
echo "== nb64+olddiag d-LU 48"; PB200_LIB=$NB64 PB200_DIAG_OLD=1 run 48 cd lu d --reps=3
