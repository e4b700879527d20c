"""GPU parity of the opt-in matrix-free operator (pf2_csr_matrix_free, SpMV variant 41): the product against the assembled CSR
product and the oracle, and whole design runs with it against the reference's golden outputs."""
import os

import numpy as np
import pytest

from oracle import portlib as orc
from pansfem2_b200 import capi, mesher, problems
from pansfem2_b200 import eqcode as ec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("make", [lambda: problems.cantilever2d(12, 8), lambda: problems.heat2d(10, 6), lambda: problems.cantilever3d(6, 4, 5),
                                  lambda: problems.cantilever2d(200, 120), lambda: problems.cantilever3d(24, 12, 10)])
def test_product_vs_csr_and_oracle(ctx, make):
    P = make()
    rng = np.random.default_rng(3)
    rho = rng.uniform(0.05, 1.0, P.nelem)
    fixed = (P.fixed[0], P.fixed[1], rng.uniform(-0.02, 0.02, len(P.fixed[0])))          # non-zero Dirichlet values: only F sees them
    mesh = capi.Mesh(ctx, P.coords, P.conn)
    dm = capi.DofMap(ctx, P.nnode, P.ndof, fixed)
    A = capi.Csr.pattern(ctx, mesh, dm)
    prm = (P.E0, P.E1, P.poisson, P.penal, P.thickness)
    A.assemble(mesh, dm, P.eq, prm, P.loads, rho=ctx.array(rho))
    x = rng.uniform(-1, 1, A.rows)
    y_csr = A.spmv_host(x)
    A.matrix_free(mesh, dm, P.eq)
    A.assemble(mesh, dm, P.eq, prm, P.loads, rho=ctx.array(rho))
    y_mf = A.spmv_host(x)
    assert rel(y_mf, y_csr) < 1e-13
    if P.nelem < 5000:
        Emod = P.E1 * rho ** P.penal + P.E0 * (1.0 - rho ** P.penal)
        So = orc.assemble(P.eq, P.coords, P.conn, fixed, P.loads, Emod, P.poisson, P.thickness)[0]
        assert rel(y_mf, So.spmv(x)) < 1e-13
    # solves agree (same preconditioner, same stopping test)
    F = A.download()[3]
    x1, it1, rr1 = A.solve_host(capi.SOLVER_SCALINGCG, F)
    A.set_spmv_variant(0)
    x0, it0, rr0 = A.solve_host(capi.SOLVER_SCALINGCG, F)
    assert rr1 < 1e-10 and abs(it1 - it0) <= max(2, it0 // 100)
    assert rel(x1, x0) < 1e-7
    for o in (A, dm, mesh):
        o.close()


def test_non_lattice_meshes_are_refused(ctx):
    P = problems.cantilever2d(12, 8)
    conn = P.conn.copy()
    conn[[3, 7]] = conn[[7, 3]]                                  # permuted element order
    coords = P.coords.copy()
    cases = [(coords, conn)]
    c2 = coords.copy(); c2[40, 0] += 0.01                       # one distorted element
    cases.append((c2, P.conn))
    for co, cn in cases:
        mesh = capi.Mesh(ctx, co, cn)
        dm = capi.DofMap(ctx, P.nnode, 2, P.fixed)
        A = capi.Csr.pattern(ctx, mesh, dm)
        with pytest.raises(capi.Pf2Error):
            A.matrix_free(mesh, dm, P.eq)
        for o in (A, dm, mesh):
            o.close()
    Pt = problems.family_problem(ec.eq_code(ec.PHYS_PLANESTRAIN, ec.SHAPE_T3), (6, 4))
    mesh = capi.Mesh(ctx, Pt.coords, Pt.conn)
    dm = capi.DofMap(ctx, Pt.nnode, 2, Pt.fixed)
    A = capi.Csr.pattern(ctx, mesh, dm)
    with pytest.raises(capi.Pf2Error):
        A.matrix_free(mesh, dm, Pt.eq)
    for o in (A, dm, mesh):
        o.close()


@pytest.mark.parametrize("opt,tag,niter", [(problems.OPT_OC, "oc", 66), (problems.OPT_MMA, "mma", 56)])
def test_simp_c1_full_run_matrix_free_vs_golden_vtk(ctx, golden_dir, opt, tag, niter):
    """The whole sample run with the matrix-free operator reproduces the committed Density_{OC,MMA}.vtk like the CSR path does."""
    g = np.load(os.path.join(golden_dir, f"density_{tag}.npz"))
    P = problems.cantilever2d(60, 40, opt_kind=opt)
    S = capi.Simp(ctx, P, matrix_free=True)
    k = 0
    for k in range(500):
        st = S.iterate(check_convergence=True)
        assert st["cg_relres"] < 1e-10
        if st["converged"]:
            break
    assert k + 1 == niter
    out = S.get(want_r=True)
    assert np.abs(out["rho"] - g["rho"]).max() < 2e-6
    np.testing.assert_allclose(out["u"], g["u"], rtol=1e-5, atol=1e-11)
    np.testing.assert_allclose(out["r"], g["r"], rtol=1e-5, atol=1e-7)
    S.close()


@pytest.mark.parametrize("make,niter", [(lambda: problems.heat2d(24, 24), 6),
                                         (lambda: problems.cantilever3d(10, 6, 4, opt_kind=problems.OPT_MMA), 5)])
def test_simp_other_configs_matrix_free_vs_oracle(ctx, make, niter):
    P = make()
    S = capi.Simp(ctx, P, matrix_free=True)
    hist = np.array([[st["f"], st["g"]] for st in (S.iterate(check_convergence=False) for _ in range(niter))])
    R = orc.simp_run(P.eq, P.coords, P.conn, P.fixed, P.loads, P.filter_kind, P.nbrs, P.opt_kind, P.optp(), P.params(), niter,
                     np.full(P.nelem, 0.5), check_convergence=False)
    np.testing.assert_allclose(hist[:, 0], R["hist"][:, 0], rtol=1e-8)
    np.testing.assert_allclose(hist[:, 1], R["hist"][:, 1], rtol=0, atol=1e-9)
    out = S.get()
    assert np.abs(out["s"] - R["s"]).max() < 1e-6
    S.close()
