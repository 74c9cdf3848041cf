mkdir -p gpurun_out
bash tools/sanitize.sh > gpurun_out/r2x_sanitize.txt 2>&1
cat gpurun_out/r2x_sanitize.txt
