set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_zz_cpp_plugin.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2g_pytest.log
cat gpurun_out/r2g_pytest.log
python tools/ef_probe.py 1e9 1.0 | tee gpurun_out/r2g_ef_probe.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ef_decode -s 3 -c 1 -o gpurun_out/r2g_ef -f python tools/ef_probe.py 1e9 1.0 > gpurun_out/r2g_ncu.log 2>&1
tail -3 gpurun_out/r2g_ncu.log
