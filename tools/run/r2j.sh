set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_blob_files.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2j_pytest.log
cat gpurun_out/r2j_pytest.log
