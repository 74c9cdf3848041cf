mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2q_pytest.log
cat gpurun_out/r2q_pytest.log
IDC_TRACE_HOST=1 python tools/rows_trace.py 2> gpurun_out/r2q_rows_trace.txt
tail -8 gpurun_out/r2q_rows_trace.txt
python tools/graph_probe.py | tee gpurun_out/r2q_graph.txt
