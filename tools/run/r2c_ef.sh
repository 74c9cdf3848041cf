set -x
mkdir -p gpurun_out
python tools/ef_probe.py 1e9 1.0 | tee gpurun_out/r2c_ef_probe.json
