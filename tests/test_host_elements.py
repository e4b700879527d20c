"""The element routines the kernels run (pansfem2_b200/csrc/element.cuh, element_generic.cuh, element_advdiff.cuh) are
__host__ __device__; tests/cpp/host_elements.cu instantiates that SAME source on the host with the library's own eq-code decoder, so
this CPU suite checks it against the live-reference fixtures and the oracle without a GPU: every <Equation, ShapeFunction,
Integration> selection, the specialised Q4 / hex8 routines, the energy forms of the sensitivity pass, and the advection-diffusion
family.  (The GPU suite repeats the comparisons through the kernels and the C ABI.)  CPU only; needs nvcc (host compilation)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import portlib as orc
from pansfem2_b200 import build as libbuild
from pansfem2_b200 import eqcode as ec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    nvcc = libbuild._nvcc()
    if shutil.which(nvcc) is None and not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    lib = libbuild.build_library()
    out = str(tmp_path_factory.mktemp("host_elements") / "libhost_elements.so")
    subprocess.run([nvcc, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-w", "-Xcompiler", "-fPIC", "-shared",
                    f"-I{ROOT}/pansfem2_b200/csrc", f"{ROOT}/tests/cpp/host_elements.cu", "-o", out,
                    f"-L{os.path.dirname(lib)}", "-lpansfem2_b200", "-Xlinker", "-rpath", "-Xlinker", os.path.dirname(lib)], check=True)
    return C.CDLL(out)


def _ptr(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def host_matrix(host, eq, xe, E, V, t, specialised=1):
    xe = np.ascontiguousarray(xe, np.float64)
    m = xe.shape[0] * ec.ndof(eq)
    Ke = np.zeros((m, m))
    assert host.pf2host_element_matrix(eq, _ptr(xe), C.c_double(E), C.c_double(V), C.c_double(t), specialised, _ptr(Ke)) == 0
    return Ke


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_every_selection_matches_the_live_reference(host, golden_dir):
    fam = np.load(os.path.join(golden_dir, "live_families.npz"))
    sel = [int(v) for v in fam["selections"]]
    assert len(sel) == 111
    for eq in sel:
        ke = host_matrix(host, eq, fam[f"xe_{eq}"], 2.5, 0.3, 0.7)
        assert rel(ke, fam[f"ke_{eq}"]) < 1e-13, ec.describe(eq)


def test_specialised_and_generic_instantiations_agree(host):
    rng = np.random.default_rng(7)
    q4 = np.array([[0, 0], [1.2, 0.1], [1.1, 0.9], [-0.1, 1.0]])
    h8 = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float) + 0.1 * rng.uniform(-1, 1, (8, 3))
    for eq, xe in ((ec.eq_code(ec.PHYS_PLANESTRAIN), q4), (ec.eq_code(ec.PHYS_HEAT), q4), (ec.eq_code(ec.PHYS_SOLID), h8)):
        a, b = host_matrix(host, eq, xe, 3.0, 0.3, 0.8, 1), host_matrix(host, eq, xe, 3.0, 0.3, 0.8, 0)
        assert rel(a, b) < 1e-14
        assert rel(a, orc.element_matrix(eq, xe, 3.0, 0.3, 0.8)) < 1e-13


def test_energy_forms_of_the_sensitivity_pass(host, golden_dir):
    """generic_energy (strain-energy form) must equal ue^T Ke ue and Ke ue of the matrix form for every selection."""
    fam = np.load(os.path.join(golden_dir, "live_families.npz"))
    rng = np.random.default_rng(11)
    for eq in (int(v) for v in fam["selections"]):
        xe = np.ascontiguousarray(fam[f"xe_{eq}"])
        nd = ec.ndof(eq)
        ue = rng.uniform(-1, 1, (xe.shape[0], nd))
        fe = np.zeros_like(ue)
        w = C.c_double()
        assert host.pf2host_element_energy(eq, _ptr(xe), _ptr(ue), C.c_double(0.3), C.c_double(0.7), C.byref(w), _ptr(fe)) == 0
        Ke = fam[f"ke_{eq}"] / 2.5          # the fixture was computed with E = 2.5
        u = ue.ravel()
        assert abs(w.value - u @ Ke @ u) < 1e-12 * np.abs(Ke).max() * (u @ u), ec.describe(eq)
        assert np.abs(fe.ravel() - Ke @ u).max() < 1e-12 * np.abs(Ke).max() * np.abs(u).max() * len(u), ec.describe(eq)


def test_advection_diffusion_family_matches_the_live_reference(host, golden_dir):
    adv = np.load(os.path.join(golden_dir, "live_advection.npz"))
    cases = adv["cases"]
    assert len(cases) == 180
    for i, (shape, quad, terms, ax, ay, k) in enumerate(cases):
        eq = ec.eq_code(ec.PHYS_ADVDIFF, int(shape), int(quad), int(terms))
        ke = host_matrix(host, eq, adv[f"xe_{i}"], ax, ay, k)
        assert rel(ke, adv[f"ke_{i}"]) < 1e-13, ec.describe(eq)


def test_advection_diffusion_groups_split(host, golden_dir):
    """advdiff_rows returns the stiffness group (A + D + AS + SC) and the mass group (M + MS) separately: the time discretisation
    of sample_advectiondiffusion_dynamic.cpp:68-69 weights them apart."""
    adv = np.load(os.path.join(golden_dir, "live_advection.npz"))
    for i, (shape, quad, terms, ax, ay, k) in enumerate(adv["cases"]):
        if int(terms) != 63:
            continue
        xe = np.ascontiguousarray(adv[f"xe_{i}"])
        n = xe.shape[0]
        KK, MM = np.zeros((n, n)), np.zeros((n, n))
        eq = ec.eq_code(ec.PHYS_ADVDIFF, int(shape), int(quad), 63)
        assert host.pf2host_advdiff_groups(eq, _ptr(xe), C.c_double(ax), C.c_double(ay), C.c_double(k), _ptr(KK), _ptr(MM)) == 0
        assert rel(KK, orc.advdiff_element(int(shape), int(quad), 15, xe, ax, ay, k)) < 1e-13
        assert rel(MM, orc.advdiff_element(int(shape), int(quad), 48, xe, ax, ay, k)) < 1e-13


def test_general_constitutive_matrix_selections_match_the_live_reference(host, golden_dir):
    """PlaneStiffness / PlaneStiffnessBbar / PlaneStiffnessWilsonTaylor (Homogenization.h:141-280): general_rows on the host."""
    pd = np.load(os.path.join(golden_dir, "live_plane_d.npz"))
    cases = [int(v) for v in pd["cases"]]
    assert len(cases) == 80
    for k, eq in enumerate(cases):
        xe, D = np.ascontiguousarray(pd[f"xe_{k}"]), np.ascontiguousarray(pd[f"D_{k}"]).reshape(9)
        m = 2 * xe.shape[0]
        Ke = np.zeros((m, m))
        assert host.pf2host_element_matrix_d(eq, _ptr(xe), _ptr(D), C.c_double(0.7), _ptr(Ke)) == 0
        assert rel(Ke, pd[f"ke_{k}"]) < 1e-12, ec.describe(eq)
        assert rel(orc.element_matrix_d(eq, xe, D, 0.7), pd[f"ke_{k}"]) == 0.0           # the restatement is bit-identical


def test_invalid_codes_are_rejected_by_the_decoder(host):
    xe, Ke = np.zeros((8, 2)), np.zeros((8, 8))
    for eq in (ec.eq_code(ec.PHYS_ADVDIFF, ec.SHAPE_T3, ec.QUAD_G1TRI, 0),       # no routine selected
               ec.eq_code(ec.PHYS_ADVDIFF, ec.SHAPE_T3, ec.QUAD_G4SQ, 1),        # square rule on a triangle
               ec.eq_code(ec.PHYS_ADVDIFF, ec.SHAPE_TET4, ec.QUAD_G1TET, 1),     # 3-D shape
               ec.eq_code(ec.PHYS_ADVDIFF, ec.SHAPE_Q4, ec.QUAD_G4SQ, 64),       # unknown routine bit
               ec.eq_code(31)):
        assert host.pf2host_element_matrix(eq, _ptr(xe), C.c_double(1.0), C.c_double(0.3), C.c_double(1.0), 1, _ptr(Ke)) == 1
