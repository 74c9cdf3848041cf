set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2m_pytest.log
cat gpurun_out/r2m_pytest.log
(time timeout 1200 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err) 2>&1 | tail -4
tail -3 gpurun_out/r2m_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2m_ncu_bench.log 2>&1
tail -2 gpurun_out/r2m_ncu_bench.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ef_ -s 12 -c 6 -o gpurun_out/r2m_ef -f python tools/ef_probe.py 1e9 1.0 > gpurun_out/r2m_ncu.log 2>&1
tail -2 gpurun_out/r2m_ncu.log
