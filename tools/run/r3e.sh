mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_blob_files.py tests/test_zz_cpp_plugin.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r3e_pytest.log
cat gpurun_out/r3e_pytest.log
python tools/ef_probe.py 1e9 1.0 | tee gpurun_out/r3e_ef_probe.json
python tools/ef_probe2.py 2>&1 | tail -7
