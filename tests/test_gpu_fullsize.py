"""Parity at BASELINE.json's full sizes through size-independent properties (the CPU oracle cannot run these sizes in a test):
symmetry of the assembled operator, rigid-body null space, every SpMV kernel family agreeing with each other, true residual of
the converged solve, compliance identities, filter partition of unity, volume-preserving OC update."""
import numpy as np
import pytest

from pansfem2_b200 import capi, problems

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def big2d(ctx):
    """The 2 M-dof headline mesh: 1000 x 1000 Q4 plane strain, random density field (seed of SURVEY.md section 8d)."""
    P = problems.cantilever2d(1000, 1000, opt_kind=problems.OPT_OC, filter_kind=problems.FILTER_DENSITY)
    S = capi.Simp(ctx, P)
    rng = np.random.Generator(np.random.MT19937(20201017))
    rho = rng.uniform(0.01, 1.0, P.nelem)
    S.A.assemble(S.mesh, S.dofmap, P.eq, (P.E0, P.E1, P.poisson, P.penal, P.thickness), P.loads, rho=ctx.array(rho))
    yield P, S, rho
    S.close()


def test_operator_is_symmetric_and_variants_agree(ctx, big2d):
    P, S, rho = big2d
    A = S.A
    assert A.rows == 2002000 and A.nnz == 35987992
    rng = np.random.default_rng(1)
    x, y = rng.uniform(-1, 1, A.rows), rng.uniform(-1, 1, A.rows)
    ref = None
    for variant in (31, 3, 12, 22):                            # SELL-32, sub-warp CSR, shared-memory stream, TMA pipeline
        A.set_spmv_variant(variant)
        Ax, Ay = A.spmv_host(x), A.spmv_host(y)
        assert abs(y @ Ax - x @ Ay) < 1e-12 * abs(y @ Ax)      # K = K^T
        if ref is None:
            ref = Ax
        else:
            assert np.abs(Ax - ref).max() < 1e-12 * np.abs(ref).max()
    A.set_spmv_variant(0)


def test_rigid_translation_is_in_the_null_space_away_from_the_clamp(ctx, big2d):
    P, S, rho = big2d
    n2g = S.dofmap.get()
    t = np.zeros(S.A.rows)
    t[n2g[n2g[:, 0] >= 0, 0]] = 1.0                            # unit x-translation of every free node
    r = S.A.spmv_host(t)
    ny = P.grid[1]
    # rows of nodes not adjacent to the clamped column x = 0 see a pure translation: zero force
    interior = n2g[2 * (ny + 1):].ravel()
    scale = np.abs(S.A.download()[2]).max()
    assert np.abs(r[interior]).max() < 1e-9 * scale
    assert np.abs(r[n2g[(ny + 1):2 * (ny + 1)].ravel()]).max() > 1e-3 * scale / 1e6   # ... and the first free column does not


def test_solve_true_residual_and_compliance_identity(ctx, big2d):
    P, S, rho = big2d
    A = S.A
    x = ctx.empty(A.rows)
    it, relres = A.solve(capi.SOLVER_SCALINGCG, A.device_F(), x)
    assert relres < 1e-10 and 1000 < it < 100000
    F = A.download()[3]
    xh = x.download()
    r = F - A.spmv_host(xh)
    assert np.linalg.norm(r) < 5e-10 * np.linalg.norm(F)       # true residual tracks the recursive one
    # compliance by the reaction route (driver :136-153) equals F.u of the reduced system
    u = ctx.empty(P.nnode * 2)
    S.dofmap.disassemble(x, u)
    f, dfdrho, _ = capi.compliance_sens(S.mesh, P.eq, u, ctx.array(rho), (P.E0, P.E1, P.poisson, P.penal, P.thickness, 1.0))
    assert abs(f - F @ xh) < 1e-8 * abs(f)
    assert (dfdrho <= 0).all()                                 # stiffer never increases compliance


def test_filter_partition_of_unity_and_oc_volume(ctx, big2d):
    P, S, rho = big2d
    flt = S.filter
    ones = np.ones(P.nelem)
    assert np.abs(flt.apply_host(ones) - 1.0).max() < 1e-14    # weights are normalised per row
    rng = np.random.default_rng(2)
    s = rng.uniform(0.2, 0.8, P.nelem)
    dfds = -rng.uniform(0.5, 2.0, P.nelem) * 1.0e-3          # puts the multiplier inside [lambdamin, lambdamax]
    dgds = np.full(P.nelem, 1.0 / (0.5 * P.nelem))
    oc = capi.OC(ctx, P.nelem, *P.oc)
    x, steps, lam = oc.update_host(flt, 0.5, 1.0, s, 1.0, dfds, dgds)
    assert 5 <= steps <= 80
    assert (x >= np.maximum(0.0, 0.85 * s) - 1e-15).all() and (x <= np.minimum(1.0, 1.15 * s) + 1e-15).all()   # move limits OC.h:87-91
    g = flt.apply_host(x).sum() / (0.5 * P.nelem) - 1.0
    assert abs(g) < 5e-3                                        # bisection stops at 1e-3 relative lambda width
    oc.close()


def test_mid_size_400x200_value_for_value_against_the_oracle(ctx):
    """160 k dof (the largest mesh the CPU oracle solves in seconds): K (pattern and values), F, u, compliance and df/drho of one design
    pass compared entry by entry with the oracle -- sample_optimize_density_oc.cpp:113-162 on a random density field."""
    from oracle import portlib as orc
    P = problems.cantilever2d(400, 200, opt_kind=problems.OPT_OC, filter_kind=problems.FILTER_DENSITY)
    rng = np.random.Generator(np.random.MT19937(20201017))
    rho = rng.uniform(0.01, 1.0, P.nelem)
    Emod = P.E1 * rho ** P.penal + P.E0 * (1.0 - rho ** P.penal)
    orc.set_num_threads(8)
    So, n2g_o, ufix_o, _ = orc.assemble(P.eq, P.coords, P.conn, P.fixed, P.loads, Emod, P.poisson, P.thickness)
    ip_o, ix_o, da_o, F_o = So.arrays()
    S = capi.Simp(ctx, P)
    rho_d = ctx.array(rho)
    S.A.assemble(S.mesh, S.dofmap, P.eq, (P.E0, P.E1, P.poisson, P.penal, P.thickness), P.loads, rho=rho_d)
    ip, ix, da, F = S.A.download()
    assert S.A.rows == 160800 and np.array_equal(ip, ip_o) and np.array_equal(ix, ix_o)           # numbering and pattern: bit-exact
    assert np.array_equal(S.dofmap.get(), n2g_o.reshape(P.nnode, P.ndof))
    assert np.abs(da - da_o).max() < 1e-13 * np.abs(da_o).max() and np.array_equal(F, F_o)
    xo, it_o, _ = So.solve(1, F_o)
    x_d = ctx.empty(S.A.rows)
    for mode in (0, 1):                                        # three-kernel loop and persistent kernel
        S.A.set_pcg_mode(mode)
        it, relres = S.A.solve(capi.SOLVER_SCALINGCG, S.A.device_F(), x_d)
        x = x_d.download()
        assert relres < 1e-10 and abs(it - it_o) <= max(3, it_o // 50), (mode, it, it_o)
        assert np.linalg.norm(F_o - So.spmv(x)) < 2e-10 * np.linalg.norm(F_o)
        assert np.abs(x - xo).max() < 1e-7 * np.abs(xo).max()                                   # two converged solves of a system with cond ~ 1e9
    u = np.zeros((P.nnode, P.ndof))
    free = n2g_o.reshape(P.nnode, P.ndof) >= 0
    u[free] = xo[n2g_o.reshape(P.nnode, P.ndof)[free]]
    f, dfdrho, _ = capi.compliance_sens(S.mesh, P.eq, ctx.array(u.ravel()), rho_d, (P.E0, P.E1, P.poisson, P.penal, P.thickness, P.scale0))
    fo, _, dfo = orc.compliance_sens(P.eq, P.coords, P.conn, u, rho, P.E0, P.E1, P.poisson, P.thickness, P.penal, P.scale0)
    assert abs(f - fo) < 1e-11 * abs(fo) and np.abs(dfdrho - dfo).max() < 1e-10 * np.abs(dfo).max()      # 80 k-term sums in another order
    # and through the GPU's own solution: compliance within the north_star tolerance
    u_g = np.zeros((P.nnode, P.ndof))
    u_g[free] = x[n2g_o.reshape(P.nnode, P.ndof)[free]]
    f_g, _, _ = capi.compliance_sens(S.mesh, P.eq, ctx.array(u_g.ravel()), rho_d, (P.E0, P.E1, P.poisson, P.penal, P.thickness, P.scale0))
    assert abs(f_g - fo) < 1e-8 * abs(fo)
    S.close()
