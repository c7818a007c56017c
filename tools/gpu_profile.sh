# ncu evidence of one step (run under gpurun; outputs in gpurun_out/).  usage: bash tools/gpu_profile.sh [tag]
TAG=${1:-r01b}
export PYTHONUNBUFFERED=1
O=gpurun_out
mkdir -p $O
# 1. every launch of ONE step of C2 with its device time (assembly, factorization, up_down)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/${TAG}_launches_c2.csv python tools/profile_step.py c2 > $O/${TAG}_launches_c2.log 2>&1
# 2. --set full of the dominant kernel (first 16 launches = the wide low levels) and of the two up_down sweeps, C2
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_gemm_scatter -c 12 \
  -o $O/${TAG}_full_gemm_scatter_c2 -f python tools/profile_step.py c2 > $O/${TAG}_full_gemm_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_dag3 -c 2 \
  -o $O/${TAG}_full_updown_c2 -f python tools/profile_step.py c2 > $O/${TAG}_full_updown_c2.log 2>&1
for f in full_gemm_scatter_c2 full_updown_c2; do
  ncu -i $O/${TAG}_$f.ncu-rep --page raw --csv > $O/${TAG}_${f}_raw.csv 2>/dev/null
done
ncu -i $O/${TAG}_full_updown_c2.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/${TAG}_full_updown_c2_source.csv.gz
rm -f $O/${TAG}_full_gemm_scatter_c2.ncu-rep $O/${TAG}_full_updown_c2.ncu-rep     # gpurun_out/ travels back only below 64 MiB
# 3. the dominant kernel on C3 (three launches from the middle of the tree)
if [ "${PROFILE_C3:-1}" = "1" ]; then
  timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_gemm_scatter -s 6 -c 3 \
    -o $O/${TAG}_full_gemm_scatter_c3 -f python tools/profile_step.py c3 > $O/${TAG}_full_gemm_c3.log 2>&1
  ncu -i $O/${TAG}_full_gemm_scatter_c3.ncu-rep --page raw --csv > $O/${TAG}_full_gemm_scatter_c3_raw.csv 2>/dev/null
  ncu -i $O/${TAG}_full_gemm_scatter_c3.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/${TAG}_full_gemm_scatter_c3_source.csv.gz
  rm -f $O/${TAG}_full_gemm_scatter_c3.ncu-rep
fi
du -sh $O; ls -la $O | tail -20
