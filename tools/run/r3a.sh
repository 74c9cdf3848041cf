mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r3a_pytest.log
cat gpurun_out/r3a_pytest.log
(time timeout 1200 python bench.py > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err) 2>&1 | tail -4
tail -3 gpurun_out/r3a_bench.err
