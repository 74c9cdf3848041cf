mkdir -p gpurun_out
# dram bytes per size-class launch of the two logical ROC kernels (second repetition: launches 9..16 of each)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_roc_encode -s 8 -c 8 -o gpurun_out/r4x_enc -f python tools/probe.py --n 1e9 --zipf 1.0 --ef 0 --reps 2 > gpurun_out/r4x_enc.log 2>&1; tail -1 gpurun_out/r4x_enc.log
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_roc_decode -s 8 -c 8 -o gpurun_out/r4x_dec -f python tools/probe.py --n 1e9 --zipf 1.0 --ef 0 --reps 2 > gpurun_out/r4x_dec.log 2>&1; tail -1 gpurun_out/r4x_dec.log
# launch list of the bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/r4x_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r4x_launch.log 2>&1; tail -c 300 gpurun_out/r4x_launch.log; wc -l gpurun_out/r4x_launches.csv
