mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r4z_bench.json 2> gpurun_out/r4z_bench.err; tail -c 300 gpurun_out/r4z_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r4z_ref.json 2> gpurun_out/r4z_ref.err; tail -c 400 gpurun_out/r4z_ref.json
