mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_blob_files.py tests/test_zz_cpp_plugin.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2w_pytest.log
cat gpurun_out/r2w_pytest.log
python tools/ef_probe.py 1e9 1.0 | tee gpurun_out/r2w_ef_probe.json
python tools/ef_probe.py 1e9 0 | tee gpurun_out/r2w_ef_probe_ctl.json
for n in 4 3; do IDC_EF_ENC_CTAS_PER_SM=$n python tools/ef_probe.py 1e9 1.0 | tee gpurun_out/r2w_ef_probe_ctas$n.json; done
