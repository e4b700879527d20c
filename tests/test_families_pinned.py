"""Pin the C restatement on the element families beyond Q4 / hex8: every
<Equation, ShapeFunction, Integration> selection against live-reference fixtures (tests/golden/live_families.npz), and the
reference's committed T3 outputs sample/heattransfer/static.vtk and sample/planestrain/result.vtk.  CPU only."""
import os

import numpy as np
import pytest

from oracle import portlib as orc
from oracle import reflib
from pansfem2_b200 import eqcode as ec
from pansfem2_b200 import problems

CASES = ["t3_heat", "t6_pstress", "q8_sri", "q8_pstrain", "q4_wt", "q4_bbar", "tet4", "hex20"]


@pytest.fixture(scope="module")
def fam(golden_dir):
    return np.load(os.path.join(golden_dir, "live_families.npz"))


@pytest.fixture(scope="module")
def t3(golden_dir):
    return np.load(os.path.join(golden_dir, "t3_samples.npz"))


def test_every_selection_element_matrix(fam):
    sel = [int(v) for v in fam["selections"]]
    assert len(sel) == 111
    for eq in sel:
        ke = orc.element_matrix(eq, fam[f"xe_{eq}"], 2.5, 0.3, 0.7)
        ref = fam[f"ke_{eq}"]
        assert ke.shape == ref.shape
        assert np.abs(ke - ref).max() <= 2e-15 * np.abs(ref).max(), ec.describe(eq)      # Q4 / hex8 are bit-exact; hex20 with 27 points: a few ulp
        assert np.abs(ke - ke.T).max() <= 1e-13 * np.abs(ke).max()


@pytest.mark.parametrize("nm", CASES)
def test_family_systems_and_simp_history(fam, nm):
    eq, n = int(fam[f"{nm}_eq"]), tuple(int(v) for v in fam[f"{nm}_n"])
    P = problems.family_problem(eq, n)
    fixed = (P.fixed[0], P.fixed[1], np.where(P.fixed[1] == 0, 0.01, -0.02))
    S, n2g, ufix, _ = orc.assemble(eq, P.coords, P.conn, fixed, P.loads, fam[f"{nm}_Emod"], 0.3, 0.8)
    indptr, indices, data, F = S.arrays()
    assert np.array_equal(indptr, fam[f"{nm}_indptr"]) and np.array_equal(indices, fam[f"{nm}_indices"])
    scale = np.abs(fam[f"{nm}_data"]).max()
    assert np.abs(data - fam[f"{nm}_data"]).max() <= 1e-14 * scale
    np.testing.assert_allclose(F, fam[f"{nm}_F"], rtol=0, atol=1e-14 * max(np.abs(fam[f"{nm}_F"]).max(), scale))
    x, it, relres = S.solve(1, F)
    assert relres < 1e-10
    np.testing.assert_allclose(x, fam[f"{nm}_x"], rtol=0, atol=1e-9 * np.abs(fam[f"{nm}_x"]).max())
    R = orc.simp_run(eq, P.coords, P.conn, P.fixed, P.loads, P.filter_kind, P.nbrs, P.opt_kind, P.optp(), P.params(), 4,
                     np.full(P.nelem, 0.5), check_convergence=False)
    np.testing.assert_allclose(R["hist"][:, 0], fam[f"{nm}_hist"][:, 0], rtol=1e-8)
    np.testing.assert_allclose(R["hist"][:, 1], fam[f"{nm}_hist"][:, 1], rtol=0, atol=1e-9)
    assert np.abs(R["s"] - fam[f"{nm}_s4"]).max() < 1e-6


def t3_heat_problem(t3):
    """sample/heattransfer/sample_heattransfer_static.cpp:34-50: T = 300 on x = 0, T = 0 on x = 1, conductivity 5, t = 1."""
    coords, conn = t3["heat_coords"], t3["heat_conn"]
    left = np.nonzero(np.abs(coords[:, 0]) < 1e-5)[0]
    right = np.nonzero(np.abs(coords[:, 0] - 1.0) < 1e-5)[0]
    fn = np.concatenate([left, right]).astype(np.int32)
    fixed = (fn, np.zeros_like(fn), np.concatenate([np.full(len(left), 300.0), np.zeros(len(right))]))
    loads = (np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0))
    return coords, conn, fixed, loads


def t3_planestrain_problem(t3):
    """sample/planestrain/sample_planestrain.cpp:23-66: 4 T3 elements, E = 210000, V = 0.3, t = 1; body force (0,-300) through
    PlaneStrainBodyForce<T3, Gauss1Triangle> (= f A / 3 per node), traction (0,-200) on edges {3,4}, {4,5} through
    PlaneStrainSurfaceForce<2Line, Gauss1Line> (= f L / 2 per node), point load -100 on (3, 1)."""
    coords, conn = t3["ps_coords"], t3["ps_conn"]
    fixed = (np.array([0, 0, 5], np.int32), np.array([0, 1, 0], np.int32), np.zeros(3))
    fy = np.zeros(len(coords))
    for el in conn:
        X = coords[el]
        area = 0.5 * ((X[0, 0] - X[2, 0]) * (X[1, 1] - X[2, 1]) - (X[0, 1] - X[2, 1]) * (X[1, 0] - X[2, 0]))
        fy[el] += -300.0 * area / 3.0
    for a, b in ((3, 4), (4, 5)):
        L = np.linalg.norm(coords[a] - coords[b])
        fy[[a, b]] += -200.0 * L / 2.0
    fy[3] += -100.0
    nodes = np.arange(len(coords), dtype=np.int32)
    loads = (nodes, np.ones_like(nodes), fy)
    return coords, conn, fixed, loads


def test_t3_heat_static_vtk(t3):
    coords, conn, fixed, loads = t3_heat_problem(t3)
    eq = ec.eq_code(ec.PHYS_HEAT, ec.SHAPE_T3, ec.QUAD_G1TRI)
    S, n2g, ufix, _ = orc.assemble(eq, coords, conn, fixed, loads, np.full(len(conn), 5.0), 0.0, 1.0)
    x, it, relres = S.solve(0, S.arrays()[3])                 # the sample calls CG
    assert relres < 1e-10
    T = np.where(n2g[:, 0] >= 0, x[np.maximum(n2g[:, 0], 0)], ufix[:, 0])
    # the VTK holds coordinates and temperatures at 6 significant digits
    np.testing.assert_allclose(T, t3["heat_T"], rtol=0, atol=300.0 * 2e-5)


def test_t3_planestrain_result_vtk(t3):
    coords, conn, fixed, loads = t3_planestrain_problem(t3)
    eq = ec.eq_code(ec.PHYS_PLANESTRAIN, ec.SHAPE_T3, ec.QUAD_G1TRI)
    S, n2g, ufix, _ = orc.assemble(eq, coords, conn, fixed, loads, np.full(len(conn), 210000.0), 0.3, 1.0)
    x, it, relres = S.solve(0, S.arrays()[3])
    u = np.where(n2g >= 0, x[np.maximum(n2g, 0)], ufix)
    np.testing.assert_allclose(u, t3["ps_u"], rtol=1e-5, atol=1e-9)
    f, r, _ = orc.compliance_sens(eq, coords, conn, u, np.ones(len(conn)), 0.0, 210000.0, 0.3, 1.0, 1.0, 1.0)
    np.testing.assert_allclose(r, t3["ps_r"], rtol=1e-5, atol=2e-3)


@pytest.mark.skipif(not reflib.available(), reason="live reference not built (needs /root/reference)")
def test_port_vs_live_reference_families_fresh():
    reflib.set_num_threads(1)
    rng = np.random.default_rng(11)
    for eq, n in ((ec.eq_code(ec.PHYS_PLANESTRESS, ec.SHAPE_Q8, ec.QUAD_G4SQ), (3, 3)), (ec.eq_code(ec.PHYS_SOLID, ec.SHAPE_HEX20), (2, 2, 2)),
                  (ec.eq_code(ec.PHYS_PLANESTRAIN_SRI, ec.SHAPE_T6, ec.QUAD_G3TRI, ec.QUAD_G1TRI), (3, 2))):
        P = problems.family_problem(eq, n)
        Emod = rng.uniform(0.5, 2.0, P.nelem)
        Sr = reflib.assemble(eq, P.coords, P.conn, P.fixed, P.loads, Emod)
        So = orc.assemble(eq, P.coords, P.conn, P.fixed, P.loads, Emod)[0]
        a, b = Sr.arrays(), So.arrays()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert np.abs(a[2] - b[2]).max() <= 1e-14 * np.abs(a[2]).max()
