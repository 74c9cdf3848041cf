for v in 0 1 7; do echo "no_defer=$v"; IDC_DEC_NO_DEFER=$v python tools/adv_probe.py 2>&1 | grep "^geo n"; done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -3
