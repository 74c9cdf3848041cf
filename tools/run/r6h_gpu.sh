#!/bin/bash
mkdir -p gpurun_out
PYTHONPATH=. timeout 16 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k_roc_.*(small|warp)' -c 12 --csv --log-file gpurun_out/r6h_launches.csv python tools/run/r6h.py > gpurun_out/r6h.log 2>&1; tail -2 gpurun_out/r6h.log; grep -c k_roc gpurun_out/r6h_launches.csv
