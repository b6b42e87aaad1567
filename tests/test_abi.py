"""CPU-side checks of the C ABI: the shared library loads and exports every symbol include/hmrecon.h declares, the
record structs have the documented sizes, and — with no GPU — creating an engine fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess
import pytest
from conftest import ROOT


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hmr_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from libhm_b200 import engine
    lib = C.CDLL(engine.LIB_PATH)
    names = _declared("hmrecon.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libhmrecon.so does not export {n}"


def test_record_struct_sizes_match_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include "hmr_records.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",sizeof(hmr_frame_hdr),sizeof(hmr_tu),'
                   'sizeof(hmr_intra),sizeof(hmr_ctu_intra_range),sizeof(hmr_pu),sizeof(hmr_sao),sizeof(hmr_ctu),sizeof(hmr_frame_desc));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert out == ["80", "20", "16", "24", "16", "10", "36", "112"]
    from libhm_b200 import records
    assert records.HDR_DT.itemsize == 80 and C.sizeof(records.FrameDesc) == 112


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from libhm_b200 import engine
    with pytest.raises(engine.EngineError):
        engine.Engine(0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under libhm_b200/ or frontend/ may reference it."""
    for base in ("libhm_b200", "frontend"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "_build" in dirpath or "_obj" in dirpath:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".c")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "hm_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, os.path.join(dirpath, f)
