set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2n_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2n_bench_n8.json 2> gpurun_out/r2n_bench_n8.err
tail -5 gpurun_out/r2n_bench_n8.err
python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2n_pytest_sharded.log
cat gpurun_out/r2n_pytest_sharded.log
