#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
run() { timeout 300 python tools/run_case.py "$@" 2>&1 | grep -E 'factorize' | tail -2; }
for M in 8 32 74 148 296; do
echo "== C2 pdlmax=$M"; PB200_PDL_MAX=$M run 64 7 llt d --reps=3
echo "== c4s pdlmax=$M"; PB200_PDL_MAX=$M run 64 cd lu z --reps=2
done
echo "== C3 pdlmax=32"; PB200_PDL_MAX=32 run 100 27 ldlt d --reps=2
echo "== C2 pdl=0"; PB200_PDL=0 run 64 7 llt d --reps=3
echo "== c4s pdl=0"; PB200_PDL=0 run 64 cd lu z --reps=2
