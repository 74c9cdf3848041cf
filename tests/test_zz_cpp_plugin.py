"""The C++ host side end to end: tests/cpp/plugin_main.cpp drives the reference's plugin classes as defined by
csrc/plugin/idc_faiss_plugin.h (on top of the C ABI, Faiss replaced by tests/faiss_shim.h) the way the reference's
tests do (test_compressed_ivfs.py:26-90, test_altid.py:19-44). CPU: it compiles, links against libidcodec.so and
fails loudly without a device. GPU: every check passes (first run on a B200: profiles/r1_plugin_cpp_b200.txt).
Named zz so that it runs after the parity suites. `plugin_main --all` adds the packed-bits classes (not yet run on a GPU)."""
import subprocess
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "cpp" / "plugin_main.cpp"
EXE = ROOT / "tests" / "cpp" / "plugin_main"


def build_plugin_main() -> Path:
    from vector_db_id_compression_b200 import build

    lib = build.build()
    deps = [SRC, ROOT / "tests" / "faiss_shim.h", ROOT / "include" / "idcodec.h",
            ROOT / "vector_db_id_compression_b200" / "csrc" / "plugin" / "idc_faiss_plugin.h", lib]
    if EXE.exists() and all(d.stat().st_mtime <= EXE.stat().st_mtime for d in deps):
        return EXE
    cc = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    subprocess.run([cc, "-std=c++17", "-O1", "-Wall", f"-I{ROOT / 'tests'}", f"-I{ROOT / 'include'}",
                    f"-I{ROOT / 'vector_db_id_compression_b200' / 'csrc' / 'plugin'}", str(SRC), "-o", str(EXE),
                    f"-L{lib.parent}", "-lidcodec", "-Wl,-rpath,$ORIGIN/../../vector_db_id_compression_b200",
                    "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    return EXE


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_cpp_plugin_links_and_refuses_to_run_without_a_device():
    exe = build_plugin_main()
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_cpp_plugin_classes_on_the_gpu():
    exe = EXE if EXE.exists() else build_plugin_main()  # built by __graft_entry__.build(); travels like the .so files
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "all checks passed" in r.stdout, (r.returncode, r.stdout, r.stderr)
