"""Host side of the header mirror (pansfem2_b200/src): the shape-function and integration policy classes must evaluate like
the reference's (they feed the host-side load vectors and user code).  tests/cpp/shape_tables.cpp is compiled against the mirror
here and its output compared with the same program built against the reference's headers (tests/golden/shape_tables.txt,
written by tests/golden/make_golden.py).  CPU only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shape_and_gauss_tables_match_the_reference(tmp_path, golden_dir):
    exe = tmp_path / "shape_tables"
    subprocess.run(["g++", "-O1", "-std=c++17", f"-I{ROOT}/pansfem2_b200/src", f"-I{ROOT}/include",
                    f"{ROOT}/tests/cpp/shape_tables.cpp", "-o", str(exe)], check=True)
    got = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().split("\n")
    want = open(os.path.join(golden_dir, "shape_tables.txt")).read().strip().split("\n")
    assert len(got) == len(want) == 103
    for lg, lw in zip(got, want):
        tg, tw = lg.split(), lw.split()
        assert len(tg) == len(tw) and tg[0] == tw[0], (lg, lw)
        for a, b in zip(tg[1:], tw[1:]):
            if a == b:
                continue
            assert abs(float(a) - float(b)) <= 4e-16 * max(1.0, abs(float(b))), (lg, lw)
