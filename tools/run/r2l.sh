mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2l_pytest.log
cat gpurun_out/r2l_pytest.log
IDC_TRACE_HOST=1 python tools/rows_trace.py 2> gpurun_out/r2l_rows_trace.txt
head -12 gpurun_out/r2l_rows_trace.txt
python tools/graph_probe.py | tee gpurun_out/r2l_graph.txt
python tools/ef_probe.py 1e9 1.0 | tee gpurun_out/r2l_ef_probe.json
