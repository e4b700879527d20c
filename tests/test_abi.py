"""CPU-side checks of the drop-in boundary: the shared library builds, loads and exports every symbol that
include/pansfem2_b200.h declares; without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from pansfem2_b200 import build
    return build.build_library()


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pansfem2_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pf2_[A-Za-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(libpath):
    lib = ctypes.CDLL(libpath)
    syms = declared_symbols()
    assert len(syms) >= 60
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/pansfem2_b200.h must compile as C99 on its own (what a cgo / JNI / ctypes-style binding would see)."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "pansfem2_b200.h"\nint main(void) { return PF2_EQ_CODE(PF2_PHYS_ADVDIFF, PF2_SHAPE_T3, PF2_QUAD_G1TRI, PF2_ADV_SUPG) == 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", f"-I{ROOT}/include", str(src)], check=True)


def test_no_cpu_fallback(libpath):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pansfem2_b200 import capi
    with pytest.raises(capi.Pf2Error) as e:
        capi.Context(0)
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under pansfem2_b200/ or include/ may reference it."""
    bad = []
    for base in ("pansfem2_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|libpf2oracle|libpf2ref", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
