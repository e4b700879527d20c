"""Host side of the header mirror: the helpers that have no device path - Area, Volume, CenterOfGravity, both ElementVector overloads,
WeakSpring (with the reference's local-column quirk), LagrangeInterpolation and its derivative (FEM/Equation/General.h),
HeatTransferSurfaceFlux (HeatTransfer.h), ShapeFunction3Line (ShapeFunction.h), SetDirichlet / SetPeriodic / RemoveBoundaryConditions
(BoundaryCondition.h), every Assembling overload, Disassembling and Renumbering (Assembling.h), and the host-side integrals of Homogenization.h (HomogenizePlaneStrainBodyForce,
...Constitutive, ...Check).  tests/cpp/host_routines.cpp is compiled against the mirror here; its output must
equal, digit for digit, what it prints when built against the reference's headers (tests/golden/host_routines.txt).  CPU only."""
import os
import subprocess

from pansfem2_b200 import build as libbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_routines_match_the_reference(tmp_path, golden_dir):
    lib = libbuild.build_library()          # CSR<T> of the mirror carries a device mirror (unused here): link the library
    exe = tmp_path / "host_routines"
    subprocess.run(["g++", "-O1", "-std=c++17", "-w", f"-I{ROOT}/pansfem2_b200/src", f"-I{ROOT}/include",
                    f"{ROOT}/tests/cpp/host_routines.cpp", "-o", str(exe), f"-L{os.path.dirname(lib)}", "-lpansfem2_b200",
                    f"-Wl,-rpath,{os.path.dirname(lib)}"], check=True)
    got = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    want = open(os.path.join(golden_dir, "host_routines.txt")).read()
    assert got.count("\n") == want.count("\n") == 100
    assert got == want
