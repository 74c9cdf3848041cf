mkdir -p gpurun_out
python -m pytest tests/test_gpu_wavelet.py -x -q 2>&1 | tail -2
python tools/wt_probe.py 1e9 > gpurun_out/r5d_wt.json 2> gpurun_out/r5d_wt.err; tail -c 1500 gpurun_out/r5d_wt.json
