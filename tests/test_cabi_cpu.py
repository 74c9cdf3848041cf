"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/idcodec.h declares,
and fails loudly (no CPU fallback) when no CUDA device is present. No compute calls here."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from vector_db_id_compression_b200 import build, _lib

    build.build()
    return _lib.load()


def test_header_and_binding_agree(lib):
    from vector_db_id_compression_b200 import _lib

    header = (ROOT / "include" / "idcodec.h").read_text()
    declared = set(re.findall(r"^\s*(?:int|uint64_t|float|const char\*)\s+(idc_\w+)\s*\(", header, flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by libidcodec.so"


def test_every_entry_point_cites_the_reference():
    header = (ROOT / "include" / "idcodec.h").read_text()
    for anchor in ("codec.cpp:123-138", "codec.cpp:140-152", "custom_invlists_impl.cpp:147-194",
                   "custom_invlists_impl.cpp:210-219", "altid_impl.cpp:153-165", "elias_fano.hpp:22-57",
                   "elias_fano.hpp:141-145", "custom_invlists_impl.cpp:292-311", "altid_impl.cpp:92-101",
                   "custom_invlists_impl.cpp:346-397", "custom_invlists_impl.cpp:377-379",
                   "custom_invlists_impl.cpp:381-392"):
        assert anchor in header, anchor


def test_library_contains_sm100a_code():
    out = __import__("subprocess").run(["cuobjdump", "-lelf", str(ROOT / "vector_db_id_compression_b200" / "libidcodec.so")],
                                       capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback(lib):
    from vector_db_id_compression_b200.capi import Context, IdcError

    with pytest.raises(IdcError) as e:
        Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = ROOT / "vector_db_id_compression_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")):
        txt = f.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt and "libref_roc" not in txt, f


def test_plugin_header_compiles(tmp_path):
    """The C++ Faiss adapter (csrc/plugin/idc_faiss_plugin.h) compiles against the Faiss shim."""
    import subprocess

    src = tmp_path / "plug.cpp"
    src.write_text('#define IDC_FAISS_SHIM\n#include "faiss_shim.h"\n#include <algorithm>\n#include <string>\n'
                   '#include "idc_faiss_plugin.h"\nint main() { return 0; }\n')
    cc = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    subprocess.run([cc, "-std=c++17", "-fsyntax-only", "-Wall", f"-I{ROOT / 'tests'}", f"-I{ROOT / 'include'}",
                    f"-I{ROOT / 'vector_db_id_compression_b200' / 'csrc' / 'plugin'}", str(src)], check=True)


def test_mt19937_constants_of_the_device_decoder(tmp_path):
    """idc_core.cuh's mt_word() holds the first outputs of std::mt19937(1234) (ANSState::stack_slice's fallback source,
    codec.h:32-40) as immediates on the device; they must be what the standard generator yields."""
    import re
    import subprocess

    src = (Path(__file__).resolve().parents[1] / "vector_db_id_compression_b200" / "csrc" / "idc_core.cuh").read_text()
    m = re.search(r"constexpr uint32_t w\[kMtWords\] = \{([^}]*)\}", src)
    assert m, "mt_word constants not found"
    have = [int(x.strip().rstrip("u")) for x in m.group(1).split(",") if x.strip()]
    prog = tmp_path / "mt.cpp"
    prog.write_text('#include <random>\n#include <cstdio>\nint main(){std::mt19937 g(1234);for(int i=0;i<8;i++)printf("%u\\n",(unsigned)g());}\n')
    exe = tmp_path / "mt"
    subprocess.run(["g++", "-O1", "-o", str(exe), str(prog)], check=True)
    want = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert have == want
