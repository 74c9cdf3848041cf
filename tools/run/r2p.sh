mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_blob_files.py -m gpu -x -q -k "ef or EliasFano" 2>&1 | tail -4 > gpurun_out/r2p_pytest.log
cat gpurun_out/r2p_pytest.log
python tools/ef_probe.py 1e9 1.0 | tee gpurun_out/r2p_ef_probe.json
python tools/ef_probe.py 1e9 0 | tee gpurun_out/r2p_ef_probe_ctl.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ef_encode -s 3 -c 1 -o gpurun_out/r2p_ef -f python tools/ef_probe.py 1e9 1.0 > gpurun_out/r2p_ncu.log 2>&1
