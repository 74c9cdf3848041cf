mkdir -p gpurun_out
python tools/adv_probe.py 2>&1 | grep "^geo"
python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_blob_files.py -x -q 2>&1 | tail -3
bash tools/ab.sh "IDC_X=0" 2>&1 | tail -1
bash tools/ab.sh "IDC_X=0" --zipf-s 0 2>&1 | tail -1
