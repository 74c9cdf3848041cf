set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_zz_cpp_plugin.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2i_pytest.log
cat gpurun_out/r2i_pytest.log
for cap in 256 384 512 768; do IDC_EF_DEC_STAGE=$cap python tools/ef_probe.py 1e9 1.0 | tee gpurun_out/r2i_ef_probe_$cap.json; done
IDC_EF_DEC_STAGE=384 python tools/ef_probe.py 1e9 0 | tee gpurun_out/r2i_ef_probe_ctl.json
