"""GPU parity of the non-symmetric Krylov family (pf2_solve with PF2_SOLVER_BICGSTAB / _BICGSTAB2 / _SCALINGBICGSTAB / _ILU0BICGSTAB;
CG.h:159-253, 357-393, 458-495) through the C ABI, against live-reference solutions (tests/golden/live_krylov.npz) and the
reference's committed sample/advection/AdvectionSUPG.vtk."""
import os

import numpy as np
import pytest

from oracle import portlib as orc
from pansfem2_b200 import capi

pytestmark = pytest.mark.gpu

KINDS = [(capi.SOLVER_BICGSTAB, "bicgstab"), (capi.SOLVER_BICGSTAB2, "bicgstab2"), (capi.SOLVER_SCALINGBICGSTAB, "scalingbicgstab"),
         (capi.SOLVER_ILU0BICGSTAB, "ilu0bicgstab")]


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def kry(golden_dir):
    return np.load(os.path.join(golden_dir, "live_krylov.npz"))


@pytest.mark.parametrize("solver,nm", KINDS)
def test_solvers_vs_live_reference_fixture(ctx, kry, solver, nm):
    A = capi.Csr.upload(ctx, kry["indptr"], kry["indices"], kry["data"])
    x, it, relres = A.solve_host(solver, kry["b"])
    _, it_o, _ = orc.system_from_csr(kry["indptr"], kry["indices"], kry["data"]).solve(solver, kry["b"])
    assert relres < 1e-10 and abs(it - it_o) <= 2             # same recurrence; reductions differ in the last bits
    ref = kry[f"x_{nm}"]
    assert np.abs(x - ref).max() < 1e-8 * np.abs(ref).max()
    A.close()


def test_nonconvergence_reports_like_reference(ctx, kry):
    A = capi.Csr.upload(ctx, kry["indptr"], kry["indices"], kry["data"])
    x, it, relres = A.solve_host(capi.SOLVER_BICGSTAB, kry["b"], itrmax=5, raise_noconv=False)
    xo, it_o, rr_o = orc.system_from_csr(kry["indptr"], kry["indices"], kry["data"]).solve(3, kry["b"], itrmax=5)
    assert it == it_o == 5 and abs(relres - rr_o) < 1e-9 * rr_o + 1e-14
    assert np.abs(x - xo).max() < 1e-10 * np.abs(xo).max()
    A.close()


def test_advection_supg_sample_vtk_on_the_device(ctx, kry):
    """The system of sample_advectiondiffusion_static.cpp solved with the device BiCGSTAB -> AdvectionSUPG.vtk."""
    A = capi.Csr.upload(ctx, kry["adv_indptr"], kry["adv_indices"], kry["adv_data"])
    for solver in (capi.SOLVER_BICGSTAB, capi.SOLVER_BICGSTAB2, capi.SOLVER_SCALINGBICGSTAB, capi.SOLVER_ILU0BICGSTAB):
        x, it, relres = A.solve_host(solver, kry["adv_F"])
        assert relres < 1e-10
        n2g = kry["adv_n2g"][:, 0]
        fixed = np.zeros(len(n2g))
        fixed[kry["adv_fix_node"]] = kry["adv_fix_val"]
        T = np.where(n2g >= 0, x[np.maximum(n2g, 0)], fixed)
        np.testing.assert_allclose(T, kry["adv_T"], rtol=6e-6, atol=2e-6)
    A.close()


def test_larger_nonsymmetric_system_vs_oracle(ctx):
    """A 2-D convection-diffusion-like operator on a 300 x 200 grid (5-point stencil, upwinded): all four solvers vs the oracle."""
    nx, ny = 300, 200
    n = nx * ny
    idx = np.arange(n).reshape(nx, ny)
    rows, cols, vals = [], [], []
    def add(r, c, v):
        rows.append(r.ravel()); cols.append(c.ravel()); vals.append(np.full(r.size, v))
    add(idx, idx, 4.4)
    add(idx[1:, :], idx[:-1, :], -1.6); add(idx[:-1, :], idx[1:, :], -0.4)
    add(idx[:, 1:], idx[:, :-1], -1.3); add(idx[:, :-1], idx[:, 1:], -0.7)
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    indptr = np.zeros(n + 1, np.int64)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr)
    b = np.random.default_rng(9).uniform(-1, 1, n)
    A = capi.Csr.upload(ctx, indptr.astype(np.int32), cols.astype(np.int32), vals)
    Ao = orc.system_from_csr(indptr.astype(np.int32), cols.astype(np.int32), vals)
    for solver in (capi.SOLVER_BICGSTAB, capi.SOLVER_BICGSTAB2, capi.SOLVER_SCALINGBICGSTAB, capi.SOLVER_ILU0BICGSTAB):
        x, it, relres = A.solve_host(solver, b)
        xo, it_o, _ = Ao.solve(solver, b)
        assert relres < 1e-10 and abs(it - it_o) <= max(3, it_o // 10)
        assert np.abs(x - xo).max() < 1e-7 * np.abs(xo).max()
    A.close()
