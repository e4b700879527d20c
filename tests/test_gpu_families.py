"""GPU parity for the element families beyond Q4 / hex8, through the C ABI.

Checkers: live-reference fixtures (tests/golden/live_families.npz: every <Equation, ShapeFunction, Integration> selection of
the reference on a distorted element; assembled systems, solves and short SIMP runs on T3 / T6 / Q8 / Tet4 / Hex20 meshes), the
oracle on fresh inputs, and the reference's committed T3 outputs (sample/heattransfer/static.vtk, sample/planestrain/result.vtk).
"""
import os

import numpy as np
import pytest

from oracle import portlib as orc
from pansfem2_b200 import capi, problems
from pansfem2_b200 import eqcode as ec
from test_families_pinned import CASES, t3_heat_problem, t3_planestrain_problem


pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def fam(golden_dir):
    return np.load(os.path.join(golden_dir, "live_families.npz"))


@pytest.fixture(scope="module")
def t3(golden_dir):
    return np.load(os.path.join(golden_dir, "t3_samples.npz"))


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_every_selection_element_matrix(ctx, fam):
    sel = [int(v) for v in fam["selections"]]
    assert len(sel) == 111
    for eq in sel:
        ke = ctx.element_matrix(eq, fam[f"xe_{eq}"], 2.5, 0.3, 0.7)
        assert rel(ke, fam[f"ke_{eq}"]) < 1e-13, ec.describe(eq)


def test_general_constitutive_matrix_selections(ctx, golden_dir):
    """PlaneStiffness / PlaneStiffnessBbar / PlaneStiffnessWilsonTaylor (Homogenization.h:141-280) through pf2_element_matrix_d: 80 live-reference
    cases (every 2-D shape / rule, every <ICV, ICD> pair, symmetric and non-symmetric D).  (Added after the last GPU slot of r01; the same
    device rows are checked on the host by tests/test_host_elements.py.)"""
    pd = np.load(os.path.join(golden_dir, "live_plane_d.npz"))
    cases = [int(v) for v in pd["cases"]]
    assert len(cases) == 80
    for k, eq in enumerate(cases):
        ke = ctx.element_matrix_d(eq, pd[f"xe_{k}"], pd[f"D_{k}"], 0.7)
        assert rel(ke, pd[f"ke_{k}"]) < 1e-12, ec.describe(eq)
    # the two per-element entry points refuse each other's selections
    with pytest.raises(capi.Pf2Error):
        ctx.element_matrix(cases[0], pd["xe_0"], 1.0)
    with pytest.raises(capi.Pf2Error):
        ctx.element_matrix_d(ec.eq_code(ec.PHYS_PLANESTRAIN, ec.SHAPE_Q4), np.zeros((4, 2)), np.eye(3))


def test_invalid_selections_are_rejected(ctx):
    xe = np.zeros((4, 2))
    for eq in (ec.eq_code(ec.PHYS_PLANESTRAIN, ec.SHAPE_T3, ec.QUAD_G4SQ),        # square rule on a triangle
               ec.eq_code(ec.PHYS_SOLID, ec.SHAPE_Q4),                             # 2-D shape in a 3-D equation
               ec.eq_code(ec.PHYS_PLANESTRESS, ec.SHAPE_Q4, ec.QUAD_G4SQ, ec.QUAD_G1SQ),  # second rule without SRI
               ec.eq_code(ec.PHYS_PLANESTRAIN_WT, ec.SHAPE_Q4, ec.QUAD_G1SQ),              # Wilson-Taylor needs off-centre points
               ec.eq_code(ec.PHYS_PLANESTRAIN_WT, ec.SHAPE_T3),                            # ... and a quadrilateral
               ec.eq_code(31)):
        with pytest.raises(capi.Pf2Error):
            ctx.element_matrix(eq, xe, 1.0)


def _system(ctx, P, fixed, Emod, V, t):
    mesh = capi.Mesh(ctx, P.coords, P.conn)
    dm = capi.DofMap(ctx, P.nnode, P.ndof, fixed)
    A = capi.Csr.pattern(ctx, mesh, dm)
    A.assemble(mesh, dm, P.eq, (0.0, 0.0, V, 1.0, t), P.loads, modulus=ctx.array(Emod))
    return mesh, dm, A


@pytest.mark.parametrize("nm", CASES)
def test_family_systems_vs_live_reference_fixture(ctx, fam, nm):
    eq, n = int(fam[f"{nm}_eq"]), tuple(int(v) for v in fam[f"{nm}_n"])
    P = problems.family_problem(eq, n)
    fixed = (P.fixed[0], P.fixed[1], np.where(P.fixed[1] == 0, 0.01, -0.02))
    mesh, dm, A = _system(ctx, P, fixed, fam[f"{nm}_Emod"], 0.3, 0.8)
    indptr, indices, data, F = A.download()
    assert np.array_equal(indptr, fam[f"{nm}_indptr"]) and np.array_equal(indices, fam[f"{nm}_indices"])
    assert rel(data, fam[f"{nm}_data"]) < 1e-13
    assert np.abs(F - fam[f"{nm}_F"]).max() <= 1e-13 * max(np.abs(fam[f"{nm}_F"]).max(), np.abs(data).max())
    x, it, relres = A.solve_host(capi.SOLVER_SCALINGCG, F)
    assert relres < 1e-10
    assert rel(x, fam[f"{nm}_x"]) < 1e-8
    for o in (A, dm, mesh):
        o.close()


@pytest.mark.parametrize("nm", CASES)
def test_family_simp_history_vs_live_reference_fixture(ctx, fam, nm):
    eq, n = int(fam[f"{nm}_eq"]), tuple(int(v) for v in fam[f"{nm}_n"])
    P = problems.family_problem(eq, n)
    S = capi.Simp(ctx, P)
    hist = []
    for k in range(4):
        st = S.iterate(check_convergence=False)
        assert st["cg_relres"] < 1e-10
        hist.append((st["f"], st["g"]))
    hist = np.array(hist)
    np.testing.assert_allclose(hist[:, 0], fam[f"{nm}_hist"][:, 0], rtol=1e-8)
    np.testing.assert_allclose(hist[:, 1], fam[f"{nm}_hist"][:, 1], rtol=0, atol=1e-9)
    assert np.abs(S.get()["s"] - fam[f"{nm}_s4"]).max() < 1e-6
    S.close()


@pytest.mark.parametrize("eq,n", [(ec.eq_code(ec.PHYS_PLANESTRAIN, ec.SHAPE_T3), (40, 24)),
                                  (ec.eq_code(ec.PHYS_PLANESTRESS, ec.SHAPE_Q8, ec.QUAD_G9SQ), (20, 12)),
                                  (ec.eq_code(ec.PHYS_PLANESTRAIN_SRI, ec.SHAPE_Q4), (24, 16)),
                                  (ec.eq_code(ec.PHYS_HEAT, ec.SHAPE_T6, ec.QUAD_G3TRI), (16, 12)),
                                  (ec.eq_code(ec.PHYS_SOLID, ec.SHAPE_TET4), (8, 5, 4)),
                                  (ec.eq_code(ec.PHYS_SOLID, ec.SHAPE_HEX20, ec.QUAD_G27CUBE), (6, 4, 3)),
                                  (ec.eq_code(ec.PHYS_SOLID, ec.SHAPE_HEX8, ec.QUAD_G27CUBE), (6, 4, 4))])
def test_family_compliance_sensitivity_and_loop_vs_oracle(ctx, eq, n):
    """Larger meshes against the oracle: reaction / compliance / sensitivity pass on a random field, then 3 design iterations
    with MMA + the density filter."""
    P = problems.family_problem(eq, n, opt_kind=problems.OPT_MMA)
    rng = np.random.default_rng(5)
    u = rng.uniform(-1, 1, (P.nnode, P.ndof)) * 1e-3
    rho = rng.uniform(0.1, 1.0, P.nelem)
    mesh = capi.Mesh(ctx, P.coords, P.conn)
    f, dfd, r = capi.compliance_sens(mesh, eq, ctx.array(u.ravel()), ctx.array(rho), (P.E0, P.E1, P.poisson, P.penal, P.thickness, P.scale0), want_r=True)
    fo, ro, dfo = orc.compliance_sens(eq, P.coords, P.conn, u, rho, P.E0, P.E1, P.poisson, P.thickness, P.penal, P.scale0)
    assert abs(f - fo) <= 1e-11 * abs(fo)
    assert rel(dfd, dfo) < 1e-11 and rel(r, ro) < 1e-11
    mesh.close()
    S = capi.Simp(ctx, P)
    hist = np.array([[st["f"], st["g"]] for st in (S.iterate(check_convergence=False) for _ in range(3))])
    R = orc.simp_run(eq, P.coords, P.conn, P.fixed, P.loads, P.filter_kind, P.nbrs, P.opt_kind, P.optp(), P.params(), 3,
                     np.full(P.nelem, 0.5), check_convergence=False)
    np.testing.assert_allclose(hist[:, 0], R["hist"][:, 0], rtol=1e-8)
    np.testing.assert_allclose(hist[:, 1], R["hist"][:, 1], rtol=0, atol=1e-9)
    out = S.get()
    assert np.abs(out["s"] - R["s"]).max() < 1e-6 and np.abs(out["rho"] - R["rho"]).max() < 1e-6
    S.close()


def test_hex20_warp_kernel_equals_the_thread_kernel_and_the_oracle(ctx, monkeypatch):
    """Hex20 assembly: one warp per element (gradients of the 27 points staged once in shared memory; default) against the
    thread-per-(element, node) kernel (PF2_HEX20_WARP=0) and the oracle, on a distorted mesh with Dirichlet values."""
    eq = ec.eq_code(ec.PHYS_SOLID, ec.SHAPE_HEX20, ec.QUAD_G27CUBE)
    P = problems.family_problem(eq, (5, 3, 4))
    rng = np.random.default_rng(11)
    coords = P.coords + rng.uniform(-0.08, 0.08, P.coords.shape)
    fixed = (P.fixed[0], P.fixed[1], np.where(P.fixed[1] == 0, 0.01, -0.02))
    Emod = rng.uniform(0.1, 2.0, P.nelem)
    mesh = capi.Mesh(ctx, coords, P.conn)
    dm = capi.DofMap(ctx, P.nnode, 3, fixed)
    A = capi.Csr.pattern(ctx, mesh, dm)
    got = {}
    for sw in ("1", "0"):
        monkeypatch.setenv("PF2_HEX20_WARP", sw)
        A.assemble(mesh, dm, eq, (0.0, 0.0, 0.3, 1.0, 1.0), P.loads, modulus=ctx.array(Emod))
        got[sw] = A.download()
    assert rel(got["1"][2], got["0"][2]) < 1e-14 and np.abs(got["1"][3] - got["0"][3]).max() <= 1e-14 * np.abs(got["0"][2]).max()
    So, *_ = orc.assemble(eq, coords, P.conn, fixed, P.loads, Emod, 0.3, 1.0)
    indptr, indices, data, F = So.arrays()
    assert np.array_equal(indices, got["1"][1]) and rel(got["1"][2], data) < 1e-13
    assert np.abs(got["1"][3] - F).max() <= 1e-13 * max(np.abs(F).max(), np.abs(data).max())
    for o in (A, dm, mesh):
        o.close()


def test_t3_heat_static_vtk_on_the_device(ctx, t3):
    """sample/heattransfer/sample_heattransfer_static.cpp (HeatTransfer<T3, Gauss1Triangle> + CG) against static.vtk."""
    coords, conn, fixed, loads = t3_heat_problem(t3)
    eq = ec.eq_code(ec.PHYS_HEAT, ec.SHAPE_T3, ec.QUAD_G1TRI)
    mesh = capi.Mesh(ctx, coords, conn)
    dm = capi.DofMap(ctx, len(coords), 1, fixed)
    A = capi.Csr.pattern(ctx, mesh, dm)
    A.assemble(mesh, dm, eq, (0.0, 0.0, 0.0, 1.0, 1.0), loads, modulus=ctx.array(np.full(len(conn), 5.0)))
    F = A.download()[3]
    x, it, relres = A.solve_host(capi.SOLVER_CG, F)
    n2g = dm.get()
    T = np.where(n2g[:, 0] >= 0, x[np.maximum(n2g[:, 0], 0)], np.where(np.abs(coords[:, 0]) < 1e-5, 300.0, 0.0))
    np.testing.assert_allclose(T, t3["heat_T"], rtol=0, atol=300.0 * 2e-5)
    for o in (A, dm, mesh):
        o.close()


def test_t3_planestrain_result_vtk_on_the_device(ctx, t3):
    """sample/planestrain/sample_planestrain.cpp (PlaneStrainStiffness<T3, Gauss1Triangle> + CG) against result.vtk: u and r."""
    coords, conn, fixed, loads = t3_planestrain_problem(t3)
    eq = ec.eq_code(ec.PHYS_PLANESTRAIN, ec.SHAPE_T3, ec.QUAD_G1TRI)
    mesh = capi.Mesh(ctx, coords, conn)
    dm = capi.DofMap(ctx, len(coords), 2, fixed)
    A = capi.Csr.pattern(ctx, mesh, dm)
    A.assemble(mesh, dm, eq, (0.0, 0.0, 0.3, 1.0, 1.0), loads, modulus=ctx.array(np.full(len(conn), 210000.0)))
    x, it, relres = A.solve_host(capi.SOLVER_CG, A.download()[3])
    n2g = dm.get()
    u = np.where(n2g >= 0, x[np.maximum(n2g, 0)], 0.0)
    np.testing.assert_allclose(u, t3["ps_u"], rtol=1e-5, atol=1e-9)
    f, dfd, r = capi.compliance_sens(mesh, eq, ctx.array(u.ravel()), ctx.array(np.ones(len(conn))), (0.0, 210000.0, 0.3, 1.0, 1.0, 1.0), want_r=True)
    np.testing.assert_allclose(r, t3["ps_r"], rtol=1e-5, atol=2e-3)
    for o in (A, dm, mesh):
        o.close()
