mkdir -p gpurun_out
python -m pytest tests/test_gpu_wavelet.py tests/test_gpu_blob_files.py tests/test_zz_cpp_plugin.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2r_pytest.log
cat gpurun_out/r2r_pytest.log
