mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "host" 2>&1 | tail -5
timeout 600 python tools/e2e_probe.py > gpurun_out/r4a_e2e.txt 2>&1; tail -8 gpurun_out/r4a_e2e.txt
IDC_NO_MILESTONES=1 timeout 600 python tools/e2e_probe.py > gpurun_out/r4a_e2e_off.txt 2>&1; tail -4 gpurun_out/r4a_e2e_off.txt
