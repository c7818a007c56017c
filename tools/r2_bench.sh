# bench lines as the driver runs them: usage tools/r2_bench.sh <tag> <ngpu> [extra bench args]
export PYTHONUNBUFFERED=1
TAG=$1; N=$2; shift 2
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --gpus 1 "$@" > gpurun_out/r2_bench_${TAG}.json 2> gpurun_out/r2_bench_${TAG}.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N "$@" > gpurun_out/r2_bench_${TAG}.json 2> gpurun_out/r2_bench_${TAG}.err
fi
echo "rc=$?"; tail -c 600 gpurun_out/r2_bench_${TAG}.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_${TAG}.json"))
    keys=("value","fact_ms","solve_ms_per_rhs","pct_fp64_peak","backward_error","ms_per_step","factor_relerr_vs_n1","parity")
    print({k:d.get(k) for k in keys}); print("e2e",d.get("e2e")); print("also",json.dumps(d.get("also"))[:1500]); print("roofline frac", (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("solve"))
    print("cpu", d.get("cpu_baseline"))
except Exception as e: print("no json", e)
PY
