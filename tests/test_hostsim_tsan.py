"""CPU: the lane-group protocol of the ROC step code (csrc/roc_group.cuh, idc_core.cuh: stream ring serviced once per step,
insert deferred into the next step, shared-memory counters) under ThreadSanitizer. The host emulation runs every lane as
a free-running thread -- lanes drift apart between two rendezvous exactly as the lanes of a group do on the device -- so a
read that is not ordered against another lane's write by a rendezvous shows up as a data race here. The warp-per-unit
coders of csrc/roc_small.cuh (32 lanes) run in the same build."""
import shutil
import subprocess
from pathlib import Path

import pytest

HERE = Path(__file__).resolve().parent / "hostsim"


def test_group_protocol_is_race_free_under_tsan(tmp_path):
    cc = shutil.which("g++")
    if cc is None:
        pytest.skip("no g++")
    exe = tmp_path / "tsan_main"
    r = subprocess.run([cc, "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-pthread", "-Wno-unknown-pragmas", "-o", str(exe),
                        str(HERE / "tsan_main.cpp"), str(HERE / "hostsim.cpp")], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer build not available: " + r.stderr[-200:])
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900, env={"TSAN_OPTIONS": "halt_on_error=0"})
    out = run.stdout + run.stderr
    lines = [l for l in run.stdout.splitlines() if l.startswith("G=")]
    assert len(lines) == 12 and all(l.endswith("roundtrip=ok") for l in lines), out[-2000:]
    wl = [l for l in run.stdout.splitlines() if l.startswith("W=32")]  # the warp-per-unit coders of roc_small.cuh
    assert len(wl) == 3 and all(l.endswith("roundtrip=ok") for l in wl), out[-2000:]
    assert "ThreadSanitizer" not in out, out[-4000:]
