"""The reference's own stand-alone unit mains that exercise host-side code of the path's boundary - src/LinearAlgebra/Models/
test_LinearAlgebra.cpp (Block / Transpose / outer, cross and inner products) and src/PrePost/Mesher/test_Mesher.cpp (SquareMesh edges, edge
selection, VTK output) - are compiled
UNMODIFIED against the header mirror (the sources are symlinked next to symlinks of the mirror's headers in a scratch tree; nothing is
copied into the repo) and must print what they print when built against the reference's headers.  They print and never assert, so the
reference build is the expectation.  (src/PrePost/Import/test_import.cpp does not compile against the reference's own ImportFromVTK.h any more
- it predates the dimension argument - so the readers are covered by tests/test_io_formats.py instead.)  CPU only; skipped where
/root/reference does not exist."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
MIRROR = os.path.join(ROOT, "pansfem2_b200", "src")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="needs the reference tree")


def mirror_tree(dst):
    """dst/src: the mirror's directory structure with one symlink per header (so that a reference source symlinked into it finds the
    MIRROR's headers through its quoted, relative includes)."""
    for dp, _, files in os.walk(MIRROR):
        rel = os.path.relpath(dp, MIRROR)
        os.makedirs(os.path.join(dst, "src", rel), exist_ok=True)
        for f in files:
            os.symlink(os.path.join(dp, f), os.path.join(dst, "src", rel, f))
    os.makedirs(os.path.join(dst, "include"), exist_ok=True)
    os.symlink(os.path.join(ROOT, "include", "pansfem2_b200.h"), os.path.join(dst, "include", "pansfem2_b200.h"))


def build_and_run(src, exe, cwd):
    subprocess.run(["g++", "-O1", "-std=c++17", "-w", src, "-o", exe], check=True)
    return subprocess.run([exe], cwd=cwd, check=True, capture_output=True, text=True).stdout


@pytest.mark.parametrize("rel,artifact", [("src/LinearAlgebra/Models/test_LinearAlgebra.cpp", None),
                                          ("src/PrePost/Mesher/test_Mesher.cpp", "result.vtk")])
def test_unmodified_reference_unit_main_on_the_mirror(tmp_path, rel, artifact):
    outs = {}
    for side in ("reference", "mirror"):
        top = tmp_path / side
        top.mkdir()
        if side == "mirror":
            mirror_tree(str(top))
            src = str(top / rel)
            os.symlink(os.path.join(REF, rel), src)
        else:
            src = os.path.join(REF, rel)
            os.makedirs(top / os.path.dirname(rel))
        cwd = top / os.path.dirname(rel)
        out = build_and_run(src, str(top / "unit_main"), str(cwd))
        outs[side] = (out, (cwd / artifact).read_bytes() if artifact else b"")
    assert len(outs["reference"][0]) > 20
    assert outs["mirror"] == outs["reference"]
