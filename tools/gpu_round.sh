set -x
export PYTHONUNBUFFERED=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dist_debug.py > gpurun_out/dist_debug.log 2>&1; echo "rc=$?"
grep "bad=" gpurun_out/dist_debug.log
timeout 600 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/pytest_dist.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_dist.log
tail -5 gpurun_out/pytest_dist.log
(
export CUDA_VISIBLE_DEVICES=0
for st in 0 3000 6000 10000; do echo "== C3 stagger $st"; PB200_STAGGER_NS=$st timeout 300 python tools/run_case.py 100 27 ldlt d 2>&1 | grep -E "factorize|solve|backward"; done
) > gpurun_out/ab_gpu0.log 2>&1 &
(
export CUDA_VISIBLE_DEVICES=1
for st in 0 3000 6000; do echo "== C2 stagger $st"; PB200_STAGGER_NS=$st timeout 300 python tools/run_case.py 64 7 llt d 2>&1 | grep -E "factorize|solve|backward"; done
echo "== C2 no inv overlap"; PB200_NO_INV_OVERLAP=1 timeout 300 python tools/run_case.py 64 7 llt d 2>&1 | grep -E "factorize|solve|backward"
echo "== C2 nb64"; PB200_LIB=$PWD/pastix_b200/lib/libpastix_b200_nb64.so timeout 300 python tools/run_case.py 64 7 llt d 2>&1 | grep -E "factorize|solve|backward"
echo "== C3 nb64"; PB200_LIB=$PWD/pastix_b200/lib/libpastix_b200_nb64.so timeout 300 python tools/run_case.py 100 27 ldlt d 2>&1 | grep -E "factorize|solve|backward"
) > gpurun_out/ab_gpu1.log 2>&1 &
wait
cat gpurun_out/ab_gpu0.log gpurun_out/ab_gpu1.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.log; echo "rc=$?"
cat gpurun_out/bench_c2_n2.json
