"""The boundary types of the path on the host side of the header mirror: every public operation of
PANSFEM2::Vector<T> / Matrix<T> (arithmetic, products, Determinant / Inverse / Cofactor up to 4 x 4, stacking, blocks, streams) and the
host behaviour of LILCSR<T> / CSR<T> (insertion, explicit zeros, get of absent entries, the per-row sort of CSR(LILCSR&), set).
tests/cpp/linalg_tables.cpp is compiled against the mirror here (linked with the library: CSR<T> carries a device mirror, unused here);
its output must equal, digit for digit, what it prints when built against the reference's headers (tests/golden/linalg_tables.txt)."""
import os
import subprocess

from pansfem2_b200 import build as libbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_vector_matrix_and_sparse_containers_match_the_reference(tmp_path, golden_dir):
    lib = libbuild.build_library()
    exe = tmp_path / "linalg_tables"
    subprocess.run(["g++", "-O1", "-std=c++17", "-w", f"-I{ROOT}/pansfem2_b200/src", f"-I{ROOT}/include", f"{ROOT}/tests/cpp/linalg_tables.cpp",
                    "-o", str(exe), f"-L{os.path.dirname(lib)}", "-lpansfem2_b200", f"-Wl,-rpath,{os.path.dirname(lib)}"], check=True)
    got = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    want = open(os.path.join(golden_dir, "linalg_tables.txt")).read()
    assert got.count("\n") == want.count("\n") == 141
    assert got == want
