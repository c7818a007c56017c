set -x
export PYTHONUNBUFFERED=1
for wl in c2 c3; do
n=8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 3 --warmup 3 --workload $wl > gpurun_out/bench_${wl}_n$n.json 2> gpurun_out/bench_${wl}_n$n.log; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_${wl}_n$n.json')); print('$wl n=$n', 'fact_ms', d['fact_ms'], 'GF', d['value'], 'solve', d['solve_ms_per_rhs'], 'berr', d['backward_error'], 'e2e', d['e2e']['value'], d['e2e']['numfact_call_ms'])"
tail -3 gpurun_out/bench_${wl}_n$n.log
done
