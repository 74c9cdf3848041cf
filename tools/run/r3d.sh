python tools/ef_probe2.py 2>&1 | tail -8
