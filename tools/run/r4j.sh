timeout 600 compute-sanitizer --tool racecheck --print-limit 6 python tools/adv_probe.py 2>&1 | grep -v "^$" | tail -30
timeout 600 compute-sanitizer --tool memcheck --print-limit 6 python tools/adv_probe.py 2>&1 | grep -v "^$" | tail -12
