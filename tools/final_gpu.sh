#!/bin/bash
# one GPU call: full parity suite, wavelet-tree probe at 1e9 ids, ncu launch list + full capture of the wavelet kernels
mkdir -p gpurun_out
timeout 100 python -m pytest tests -x -q -m gpu > gpurun_out/w6_pytest.log 2>&1; tail -4 gpurun_out/w6_pytest.log
timeout 60 python tools/wt_probe.py 1e9 65536 > gpurun_out/w6_probe_1e9.json 2> gpurun_out/w6_probe.err; cat gpurun_out/w6_probe_1e9.json; tail -3 gpurun_out/w6_probe.err
WT_PROBE_REPS=1 timeout 90 ncu --set full --clock-control none --import-source on -k 'regex:k_wt_(distribute|apply|level_bits|level_scatter|replay|emit|select)' -c 10 -o gpurun_out/w6_wt -f python tools/wt_probe.py 5e7 4 > gpurun_out/w6_ncu.log 2>&1; tail -3 gpurun_out/w6_ncu.log
WT_PROBE_REPS=1 timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_wt --csv --log-file gpurun_out/w6_launches.csv python tools/wt_probe.py 1e8 65536 > /dev/null 2>&1
ls -la gpurun_out/w6_*
