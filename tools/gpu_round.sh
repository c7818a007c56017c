set -x
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_dist.py -m gpu -q > gpurun_out/pytest_dist.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_dist.log
tail -5 gpurun_out/pytest_dist.log
n=4
for wl in c3 c2; do
for mode in early noearly; do
if [ $mode = noearly ]; then export PB200_NO_EARLY_GATHER=1; else unset PB200_NO_EARLY_GATHER; fi
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 3 --warmup 3 --workload $wl > gpurun_out/bench_${wl}_n${n}_$mode.json 2> gpurun_out/bench_${wl}_n${n}_$mode.log; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_${wl}_n${n}_$mode.json')); print('$wl n=$n $mode', 'fact_ms', d['fact_ms'], 'GF', d['value'], 'berr', d['backward_error'])"
done; done
