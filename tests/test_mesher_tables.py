"""Host side of the header mirror: SquareMesh<T> (Q4 and the 8-node "2" variants, boundary edges, element / edge selection, fixed lists)
SquareMesh2<T> (graded) and the ring-shaped AnnulusMesh<T>, SquareAnnulusMesh<T> (with the reference's swapped-rectangle quirk in
GenerateFixedlist), SquareAnnulusMesh2<T>, SquareCircleAnnulusMesh<T> must generate exactly what the reference's PrePost/Mesher headers generate - same
coordinates to the last bit, same numbering, same order.  tests/cpp/mesher_tables.cpp is compiled against the mirror here and compared with its output when
built against the reference's headers (tests/golden/mesher_tables.txt; tests/golden/make_golden.py meshers).  CPU only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_square_meshers_match_the_reference(tmp_path, golden_dir):
    exe = tmp_path / "mesher_tables"
    subprocess.run(["g++", "-O1", "-std=c++17", "-w", f"-I{ROOT}/pansfem2_b200/src", f"-I{ROOT}/include",
                    f"{ROOT}/tests/cpp/mesher_tables.cpp", "-o", str(exe)], check=True)
    got = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    want = open(os.path.join(golden_dir, "mesher_tables.txt")).read()
    assert got.count("\n") == want.count("\n") == 1062
    assert got == want
