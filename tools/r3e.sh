#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_diag_blk -s 150 -c 2 \
  -o $O/r3e_full_diag_c2 -f python tools/profile_step.py c2 > $O/r3e_full_diag_c2.log 2>&1
ncu -i $O/r3e_full_diag_c2.ncu-rep --page raw --csv > $O/r3e_full_diag_c2_raw.csv 2>/dev/null
ncu -i $O/r3e_full_diag_c2.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/r3e_full_diag_c2_source.csv.gz
rm -f $O/r3e_full_diag_c2.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_trsm_mma -s 150 -c 2 \
  -o $O/r3e_full_trsm_c2 -f python tools/profile_step.py c2 > $O/r3e_full_trsm_c2.log 2>&1
ncu -i $O/r3e_full_trsm_c2.ncu-rep --page raw --csv > $O/r3e_full_trsm_c2_raw.csv 2>/dev/null
ncu -i $O/r3e_full_trsm_c2.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/r3e_full_trsm_c2_source.csv.gz
rm -f $O/r3e_full_trsm_c2.ncu-rep
ls -la $O | tail -8
