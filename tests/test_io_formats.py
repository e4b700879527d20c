"""The data formats either side of the path, on the host side of the header mirror: every writer of PrePost/Export/ExportToVTK.h, the
CSV readers of PrePost/Import/ImportFromCSV.h (nodes, elements, Dirichlet / Neumann with "free" tokens, initial values, periodic
pairs, the open-error convention) and both VTK readers (ImportFromVTK.h: the reference's forward-scanner semantics; ImportFromVTK2.h).
tests/cpp/io_formats.cpp is compiled against the mirror here; its VTK bytes and parsed values must equal what the same program produced
when built against the reference's headers (tests/golden/io_formats*.txt, io_formats_out.vtk; tests/golden/make_golden.py io).  CPU only."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tag,flags", [("", []), ("_vtk2", ["-DIO_VTK2"])])
def test_mirror_io_equals_reference_io(tmp_path, golden_dir, tag, flags):
    exe = tmp_path / "io_formats"
    subprocess.run(["g++", "-O1", "-std=c++17", "-w", *flags, f"-I{ROOT}/pansfem2_b200/src", f"-I{ROOT}/include",
                    f"{ROOT}/tests/cpp/io_formats.cpp", "-o", str(exe)], check=True)
    (tmp_path / "work").mkdir()
    got = subprocess.run([str(exe), "work"], cwd=tmp_path, check=True, capture_output=True, text=True).stdout
    want = open(os.path.join(golden_dir, f"io_formats{tag}.txt")).read()
    assert got == want
    assert (tmp_path / "work" / "out.vtk").read_bytes() == open(os.path.join(golden_dir, "io_formats_out.vtk"), "rb").read()
    if not tag:
        assert "vtk w first 0\nvtk w second 6" in got         # the forward scanner: a miss rewinds and returns an empty container
