"""GPU tests of the persistent PCG kernel (csrc/pcg_persistent.cuh) and of the warm-start overload pf2_solve_x0.

CG.h:124-154 / 420-453 restated as one cooperative kernel per solve: same recurrences and stopping rule as the three-kernel loop
and as the oracle, so solutions agree to the solver tolerance and iteration counts up to rounding of the dot products.
"""
import numpy as np
import pytest

from oracle import portlib as orc
from pansfem2_b200 import capi, problems

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def _system(P, seed=5):
    rng = np.random.default_rng(seed)
    Emod = P.E1 * rng.uniform(0.05, 1.0, P.nelem) ** 3
    So, n2g, ufix, _ = orc.assemble(P.eq, P.coords, P.conn, P.fixed, P.loads, Emod)
    return So, So.arrays()


@pytest.mark.parametrize("make", [lambda: problems.cantilever2d(60, 40), lambda: problems.heat2d(40, 30),
                                  lambda: problems.cantilever3d(10, 6, 4), lambda: problems.cantilever2d(7, 6)])
@pytest.mark.parametrize("solver", [capi.SOLVER_CG, capi.SOLVER_SCALINGCG])
def test_persistent_kernel_vs_three_kernel_loop_and_oracle(ctx, make, solver):
    P = make()
    So, (indptr, indices, data, F) = _system(P)
    xo, it_o, _ = So.solve(solver, F)
    out = {}
    for mode in (1, 0):
        A = capi.Csr.upload(ctx, indptr, indices, data)
        A.set_pcg_mode(mode)
        try:
            x, it, relres = A.solve_host(solver, F)
        except capi.Pf2Error as e:
            raise AssertionError(f"pcg mode {mode}: {e}; {A.pcg_stats()}")
        st = A.pcg_stats()
        assert st["solves"] == (1 if (mode == 1 and A.rows >= 1000) else st["solves"] if mode == 1 else 0), (mode, st)   # the path under test really ran
        assert relres < 1e-10
        assert np.linalg.norm(F - So.spmv(x)) < 2e-10 * np.linalg.norm(F)           # true residual
        assert np.abs(x - xo).max() < 1e-8 * np.abs(xo).max()
        assert abs(it - it_o) <= max(2, it_o // 50), (mode, it, it_o)
        out[mode] = (x, it)
        A.close()
    assert np.abs(out[0][0] - out[1][0]).max() < 1e-8 * np.abs(xo).max()


def test_persistent_kernel_on_the_pattern_built_matrix_with_block_deltas(ctx):
    """hex8 matrices built by pf2_csr_pattern carry one 16-bit node delta per run of 3 columns (NB = 3 instantiation)."""
    P = problems.cantilever3d(12, 6, 5)
    rng = np.random.default_rng(2)
    Emod = rng.uniform(0.5, 2.0, P.nelem)
    mesh = capi.Mesh(ctx, P.coords, P.conn)
    dm = capi.DofMap(ctx, P.nnode, P.ndof, P.fixed)
    xs = {}
    for mode in (1, 0):
        A = capi.Csr.pattern(ctx, mesh, dm)
        A.assemble(mesh, dm, P.eq, (0.0, 0.0, 0.3, 1.0, 1.0), P.loads, modulus=ctx.array(Emod))
        A.set_pcg_mode(mode)
        F = A.download()[3]
        x, it, relres = A.solve_host(capi.SOLVER_SCALINGCG, F)
        assert relres < 1e-10 and A.pcg_stats()["solves"] == mode
        xs[mode] = (x, it)
        A.close()
    assert abs(xs[0][1] - xs[1][1]) <= 3
    assert np.abs(xs[0][0] - xs[1][0]).max() < 1e-8 * np.abs(xs[0][0]).max()
    So, n2g, ufix, _ = orc.assemble(P.eq, P.coords, P.conn, P.fixed, P.loads, Emod)
    xo, _, _ = So.solve(1, So.arrays()[3])
    assert np.abs(xs[1][0] - xo).max() < 1e-8 * np.abs(xo).max()


@pytest.mark.parametrize("mode", [1, 0])
def test_iteration_cap_returns_the_last_iterate_like_the_reference(ctx, mode):
    P = problems.cantilever2d(30, 20)
    So, (indptr, indices, data, F) = _system(P)
    A = capi.Csr.upload(ctx, indptr, indices, data)
    A.set_pcg_mode(mode)
    x, it, _ = A.solve_host(capi.SOLVER_SCALINGCG, F, itrmax=7, raise_noconv=False)
    xo, ito, _ = So.solve(1, F, itrmax=7)
    assert it == 7 and np.abs(x - xo).max() < 1e-10 * np.abs(xo).max()
    A.close()


@pytest.mark.parametrize("mode", [1, 0])
def test_warm_start_same_solution_fewer_iterations(ctx, mode):
    """pf2_solve_x0: x0 = the solution of a nearby system (what a design iteration hands to the next one)."""
    P = problems.cantilever2d(60, 40)
    So, (indptr, indices, data, F) = _system(P, seed=5)
    rng = np.random.default_rng(9)
    data2 = data * (1.0 + 1e-3 * rng.standard_normal(1)[0])          # a slightly stiffer structure, same pattern
    So2 = orc.system_from_csr(indptr, indices, data2)
    xo2, it_cold, _ = So2.solve(1, F)
    A = capi.Csr.upload(ctx, indptr, indices, data)
    A.set_pcg_mode(mode)
    b = ctx.array(F)
    x = ctx.empty(A.rows)
    it0, rr0 = A.solve(capi.SOLVER_SCALINGCG, b, x)
    A.close()
    A2 = capi.Csr.upload(ctx, indptr, indices, data2)
    A2.set_pcg_mode(mode)
    it1, rr1 = A2.solve(capi.SOLVER_SCALINGCG, b, x, warm=True)
    xw = x.download()
    assert rr1 < 1e-10 and it1 < it_cold, (it1, it_cold)
    assert np.linalg.norm(F - So2.spmv(xw)) < 2e-10 * np.linalg.norm(F)
    assert np.abs(xw - xo2).max() < 1e-8 * np.abs(xo2).max()
    # starting from the converged solution itself: the stopping rule holds before the first iteration
    it2, rr2 = A2.solve(capi.SOLVER_SCALINGCG, b, x, warm=True)
    assert it2 <= 1 and rr2 < 1e-10
    A2.close()


@pytest.mark.parametrize("matrix_free", [False, True])
def test_design_loop_with_warm_start_matches_the_oracle(ctx, matrix_free):
    """rho after 12 design iterations within 1e-6 of the oracle, compliance 1e-8 relative (north_star tolerances), fewer CG iterations;
    also with K applied matrix-free (nodal-numbering PCG, whose set-up forms r0 = b - K x0 through the fused operator)."""
    P = problems.cantilever2d(60, 40, opt_kind=problems.OPT_MMA, filter_kind=problems.FILTER_DENSITY)
    R = orc.simp_run(P.eq, P.coords, P.conn, P.fixed, P.loads, P.filter_kind, P.nbrs, P.opt_kind, P.optp(), P.params(), 12,
                     np.full(P.nelem, P.s0), check_convergence=False)
    tot = {}
    for warm in (False, True):
        S = capi.Simp(ctx, P, matrix_free=matrix_free)
        S.A.set_pcg_mode(1 if warm else -1)          # the warm run also exercises the persistent kernel's x0 set-up (assembled operator)
        S.set_warm_start(warm)
        st = [S.iterate(check_convergence=False) for _ in range(12)]
        out = S.get()
        for k in range(12):
            assert abs(st[k]["f"] - R["hist"][k, 0]) < 1e-8 * abs(R["hist"][k, 0]), (warm, k)
            assert st[k]["cg_relres"] < 1e-10
        assert np.abs(out["rho"] - R["rho"]).max() < 1e-6 and np.abs(out["s"] - R["s"]).max() < 1e-6
        tot[warm] = sum(s["cg_iters"] for s in st)
        # reset: the same run again reproduces the same history (MMA asymptotes, beta schedule, warm-start state)
        S.reset()
        st2 = [S.iterate(check_convergence=False) for _ in range(3)]
        for k in range(3):
            assert abs(st2[k]["f"] - st[k]["f"]) < 1e-9 * abs(st[k]["f"]) and st2[k]["k"] == k
        S.close()
    assert tot[True] < tot[False], tot


def test_ilu0_one_launch_sweeps_equal_level_launches_and_scale(ctx, monkeypatch):
    """PreILU0 (CG.h:289-315) as one launch per sweep (one CTA walking the levels; ready words per row) against one launch per dependency
    level, on a mesh with ~600 levels; the level schedule is built on the device.  ILU0CG then needs fewer iterations than ScalingCG."""
    P = problems.cantilever2d(200, 100)
    rng = np.random.default_rng(4)
    Emod = P.E1 * rng.uniform(0.05, 1.0, P.nelem) ** 3
    mesh = capi.Mesh(ctx, P.coords, P.conn)
    dm = capi.DofMap(ctx, P.nnode, P.ndof, P.fixed)
    A = capi.Csr.pattern(ctx, mesh, dm)
    A.assemble(mesh, dm, P.eq, (0.0, 0.0, 0.3, 1.0, 1.0), P.loads, modulus=ctx.array(Emod))
    F = A.download()[3]
    b = rng.standard_normal(A.rows)
    ys = {}
    for mode in ("level", "syncfree", "cta"):                   # launch per level | ready word per row | one CTA walks the levels (default)
        monkeypatch.setenv("PF2_ILU_SWEEP", mode)
        ys[mode] = A.ilu0_solve_host(b)
    monkeypatch.delenv("PF2_ILU_SWEEP")
    assert np.array_equal(ys["level"], ys["syncfree"]) and np.array_equal(ys["level"], ys["cta"])      # same operations per row: bit-identical
    x_j, it_j, rr_j = A.solve_host(capi.SOLVER_SCALINGCG, F)
    x_i, it_i, rr_i = A.solve_host(capi.SOLVER_ILU0CG, F)
    assert rr_i < 1e-10 and it_i < it_j / 2, (it_i, it_j)
    assert np.abs(x_i - x_j).max() < 1e-7 * np.abs(x_j).max()
    A.close()
