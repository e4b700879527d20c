"""Pin the C restatement (oracle/pf2_oracle.c) before anything trusts it:
  * against the reference's own golden outputs (Density_OC.vtk, Density_MMA.vtk, result_linear.vtk),
  * against its MMA known-answer mains (test_MMA*.cpp stdout),
  * against fixtures produced by the live reference (tests/golden/live_reference.npz), and - where
    oracle/_ref/libpf2ref.so is present - against the live reference itself on fresh inputs.
CPU only; whole file runs in well under a minute.
"""
import os

import numpy as np
import pytest

from oracle import portlib as orc
from oracle import reflib
from pansfem2_b200 import problems


def sig6(a):
    """Round to the 6 significant digits the reference's VTK writer prints (ExportToVTK.h:34,112,135)."""
    a = np.asarray(a, dtype=np.float64)
    return np.array([float("%.6g" % v) for v in a.ravel()]).reshape(a.shape)


@pytest.fixture(scope="module")
def live(golden_dir):
    return np.load(os.path.join(golden_dir, "live_reference.npz"))


def test_element_matrices_vs_live_fixture(live):
    cases = [("ke_ps_q4", orc.EQ_PLANESTRAIN, "q4", 1.0, 0.3, 1.0), ("ke_ps_q4d", orc.EQ_PLANESTRAIN, "q4d", 2.5, 0.3, 0.7),
             ("ke_heat_q4", orc.EQ_HEAT, "q4", 1.0, 0.0, 1.0), ("ke_heat_q4d", orc.EQ_HEAT, "q4d", 2.5, 0.0, 0.7),
             ("ke_solid_h8", orc.EQ_SOLID, "h8", 1.0, 0.3, 1.0), ("ke_solid_h8d", orc.EQ_SOLID, "h8d", 2.5, 0.3, 1.0)]
    for key, eq, xkey, E, V, t in cases:
        Ke = orc.element_matrix(eq, live[xkey], E, V, t)
        assert np.array_equal(Ke, live[key]), key      # same operations in the same order: bit-identical


def test_known_answers_survey_appendix_c():
    q4 = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], float)
    Ke = orc.element_matrix(orc.EQ_PLANESTRAIN, q4, 1.0, 0.3, 1.0)
    np.testing.assert_allclose(Ke[0], [0.57692307692307687, 0.24038461538461536, -0.38461538461538453, 0.048076923076923087,
                                       -0.28846153846153838, -0.24038461538461534, 0.096153846153846104, -0.048076923076923087], rtol=1e-15)
    Kh = orc.element_matrix(orc.EQ_HEAT, q4, 1.0, 0.0, 1.0)
    np.testing.assert_allclose(Kh[0], [2 / 3, -1 / 6, -1 / 3, -1 / 6], rtol=1e-15)
    h8 = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    Ks = orc.element_matrix(orc.EQ_SOLID, h8, 1.0, 0.3, 1.0)
    np.testing.assert_allclose(Ks[0, :6], [0.23504273504273504, 0.080128205128205121, 0.080128205128205121,
                                           -0.10683760683760683, 0.016025641025641024, 0.016025641025641024], rtol=1e-14)
    assert abs(np.trace(Ks) - 5.6410256410256414) < 1e-13


def test_assembly_and_solvers_vs_live_fixture(live):
    P = problems.cantilever2d(12, 8)
    fixed = (P.fixed[0], P.fixed[1], live["sys_fixval"])
    S, n2g, ufix, _ = orc.assemble(P.eq, P.coords, P.conn, fixed, P.loads, live["sys_Emod"])
    indptr, indices, data, F = S.arrays()
    assert np.array_equal(indptr, live["sys_indptr"]) and np.array_equal(indices, live["sys_indices"])
    assert np.array_equal(data, live["sys_data"])
    assert np.array_equal(F, live["sys_F"])
    for kind, nm in ((0, "cg"), (1, "scalingcg"), (2, "ilu0cg")):
        x, it, relres = S.solve(kind, F)
        assert relres < 1e-10
        np.testing.assert_allclose(x, live[f"sys_x_{nm}"], rtol=0, atol=1e-15 * np.abs(x).max() * 100)
    M = S.ilu0()
    np.testing.assert_allclose(M.arrays()[2], live["sys_ilu0_data"], rtol=1e-14, atol=0)
    np.testing.assert_allclose(M.preilu0(F), live["sys_preilu0"], rtol=1e-13, atol=1e-18)


def test_filters_and_oc_vs_live_fixture(live):
    P = problems.cantilever2d(12, 8)
    s, dfdrho = live["flt_s"], live["flt_dfdrho"]
    for kind, nm in ((orc.FILTER_DENSITY, "density"), (orc.FILTER_HEAVISIDE, "heaviside")):
        assert np.array_equal(orc.filter_apply(kind, P.nbrs, 2.0, s), live[f"flt_{nm}_rho"])
        np.testing.assert_allclose(orc.filter_sens(kind, P.nbrs, 2.0, s, dfdrho), live[f"flt_{nm}_sens"], rtol=1e-15)
        x, steps, lam = orc.oc_update(P.oc, kind, P.nbrs, 2.0, 0.5, 1.0, s, live[f"flt_{nm}_sens"], live[f"oc_{nm}_dgds"])
        assert np.array_equal(x, live[f"oc_{nm}_x"])
        assert 10 <= steps <= 60


def test_mma_known_answer_mains(golden_dir):
    kat = np.load(os.path.join(golden_dir, "mma_kat.npz"))
    # test_MMA.cpp: Svanberg 5-bar cantilever (n=5, m=1)
    C1, C2 = 0.0624, 1.0
    coef = np.array([61.0, 37.0, 19.0, 7.0, 1.0])
    mma = orc.MMA(5, 1, 1.0, [0.0], [1000.0], [1.0], 1.0, 10.0)
    mma.set_parameters(1.0e-5, 0.1, 0.5, 0.5, 0.7, 1.2)
    s = np.full(5, 5.0)
    rows = kat["test_MMA"]
    fprev = 0.0
    for k in range(len(rows)):
        f = C1 * s.sum()
        g = (coef * s ** -3.0).sum() - C2
        np.testing.assert_allclose(sig6(np.concatenate([[f, g], s])), rows[k, 1:], rtol=2e-6, atol=2e-12)
        if orc.is_convergence(f, fprev, 1e-5):
            break
        s = mma.update(s, np.full(5, C1), [g], (-3.0 * coef * s ** -4.0)[None, :])
        fprev = f
    assert k == len(rows) - 1 and abs(f - 1.33996) < 1e-5
    # test_MMA_3.cpp (two-bar truss, n = m = 2 -> the n <= m branch MMA.h:292) final answer
    assert abs(kat["test_MMA_3"][-1, 1] - 1.50865) < 1e-5


def test_mma_two_bar_truss_n_le_m_branch(golden_dir):
    """test_MMA_3.cpp restated: n = 2, m = 2 exercises MMA.h:292-330."""
    kat = np.load(os.path.join(golden_dir, "mma_kat.npz"))["test_MMA_3"]
    mma = orc.MMA(2, 2, 1.0, [0.0, 0.0], [1000.0, 1000.0], [1.0, 1.0], np.array([0.2, 0.1]), np.array([4.0, 1.6]))
    mma.set_parameters(1.0e-5, 0.1, 0.5, 0.5, 0.7, 1.2)
    s = np.array([1.5, 0.5])
    c1, c2 = 1.0, 0.124
    fprev, f = 0.0, 0.0
    for k in range(20):
        x1, x2 = s
        rt = np.sqrt(1.0 + x2 * x2)
        f = c1 * x1 * rt
        df = np.array([c1 * rt, c1 * x1 * x2 / rt])
        g1 = c2 * rt * (8.0 / x1 + 1.0 / (x1 * x2)) - 1.0
        g2 = c2 * rt * (8.0 / x1 - 1.0 / (x1 * x2)) - 1.0
        dg1 = np.array([c2 * rt * (-8.0 / x1 ** 2 - 1.0 / (x1 ** 2 * x2)), c2 * (x2 / rt * (8.0 / x1 + 1.0 / (x1 * x2)) - rt / (x1 * x2 ** 2))])
        dg2 = np.array([c2 * rt * (-8.0 / x1 ** 2 + 1.0 / (x1 ** 2 * x2)), c2 * (x2 / rt * (8.0 / x1 - 1.0 / (x1 * x2)) + rt / (x1 * x2 ** 2))])
        if orc.is_convergence(f, fprev, 1e-5):
            break
        s = mma.update(s, df, [g1, g2], np.stack([dg1, dg2]))
        fprev = f
    # Svanberg's two-bar truss optimum, and the reference's printed final objective
    assert abs(f - 1.50865) < 2e-4 and abs(kat[-1, 1] - 1.50865) < 1e-5
    np.testing.assert_allclose(s, [1.41171, 0.376912], rtol=2e-3)


@pytest.mark.parametrize("tag,opt", [("oc", problems.OPT_OC), ("mma", problems.OPT_MMA)])
def test_c1_history_vs_live_fixture(live, tag, opt):
    P = problems.cantilever2d(60, 40, opt_kind=opt)
    R = orc.simp_run(P.eq, P.coords, P.conn, P.fixed, P.loads, P.filter_kind, P.nbrs, P.opt_kind, P.optp(), P.params(),
                     12, np.full(P.nelem, 0.5), check_convergence=False)
    np.testing.assert_allclose(R["hist"][:, 0], live[f"c1_{tag}_hist"][:, 0], rtol=1e-8)       # compliance 1e-8 rel
    np.testing.assert_allclose(R["hist"][:, 1], live[f"c1_{tag}_hist"][:, 1], rtol=0, atol=1e-9)
    assert np.abs(R["rho"] - live[f"c1_{tag}_rho12"]).max() < 1e-6                              # density 1e-6 max-abs
    assert np.abs(R["s"] - live[f"c1_{tag}_s12"]).max() < 1e-6


def test_c1_oc_full_run_vs_density_oc_vtk(golden_dir):
    """The whole 66-iteration sample run reproduces the committed Density_OC.vtk at its printed precision."""
    g = np.load(os.path.join(golden_dir, "density_oc.npz"))
    P = problems.cantilever2d(60, 40)
    R = orc.simp_run(P.eq, P.coords, P.conn, P.fixed, P.loads, P.filter_kind, P.nbrs, P.opt_kind, P.optp(), P.params(),
                     500, np.full(P.nelem, 0.5))
    assert R["iters"] == 66
    assert abs(R["hist"][-1, 0] / P.scale0 - 1.51897e-4) < 1e-9
    assert np.abs(R["rho"] - g["rho"]).max() < 2e-6
    # goldens carry 6 significant digits: half a unit in the 6th digit is 5e-6 relative
    np.testing.assert_allclose(R["u"], g["u"], rtol=6e-6, atol=1e-12)
    np.testing.assert_allclose(R["r"], g["r"], rtol=6e-6, atol=1e-9)


@pytest.fixture(scope="module")
def live_conlin(golden_dir):
    return np.load(os.path.join(golden_dir, "live_conlin.npz"))


def test_sensitivity_filters_vs_live_fixture(live_conlin):
    """SensitivityFilter (Sigmund) / SensitivityFilter2 (Borrvall), SensitivityFilter.h:44-55, 88-99: bit-exact."""
    P = problems.cantilever2d(12, 8)
    for kind, nm in ((2, "sigmund"), (3, "borrvall")):
        out = orc.sensitivity_filter(kind, P.nbrs, live_conlin["s"], live_conlin["dfds"])
        assert np.array_equal(out, live_conlin[f"sens_{nm}"])


def _conlin_inputs(n, m, it, x, dfds):
    df = dfds * (1.0 + 0.1 * it) * np.where(np.arange(n) % 7 == 3, -0.05, 1.0)
    g = np.array([x.sum() / (0.5 * n) - 1.0, 0.3 - x[: n // 2].sum() / n][:m])
    dg = np.stack([np.full(n, 1.0 / (0.5 * n)), np.where(np.arange(n) < n // 2, -1.0 / n, 0.0)][:m])
    return df, g, dg


@pytest.mark.parametrize("m", [1, 2])
def test_conlin_updates_vs_live_fixture(live_conlin, m):
    """CONLIN<T>::UpdateVariables (CONLIN.h:89-373) with one and two constraints and mixed-sign gradients."""
    n = 96
    opt = orc.MMA(n, m, 1.0, np.zeros(m), np.full(m, 1.0e4), np.zeros(m), 0.01, 1.0)
    opt.set_conlin(0.2)
    x = live_conlin["s"].copy()
    for it in range(3):
        df, g, dg = _conlin_inputs(n, m, it, x, live_conlin["dfds"])
        x = opt.update(x, df, g, dg)
        assert np.abs(x - live_conlin[f"conlin_m{m}_x"][it]).max() < 1e-12


def test_c1_conlin_history_vs_live_fixture(live_conlin):
    P = problems.cantilever2d(60, 40, opt_kind=problems.OPT_CONLIN)
    R = orc.simp_run(P.eq, P.coords, P.conn, P.fixed, P.loads, P.filter_kind, P.nbrs, P.opt_kind, P.optp(), P.params(),
                     12, np.full(P.nelem, 0.5), check_convergence=False)
    np.testing.assert_allclose(R["hist"][:, 0], live_conlin["c1_conlin_hist"][:, 0], rtol=1e-8)
    np.testing.assert_allclose(R["hist"][:, 1], live_conlin["c1_conlin_hist"][:, 1], rtol=0, atol=1e-9)
    assert np.abs(R["rho"] - live_conlin["c1_conlin_rho12"]).max() < 1e-6
    assert np.abs(R["s"] - live_conlin["c1_conlin_s12"]).max() < 1e-6


def test_c1_conlin_full_run_vs_density_conlin_vtk(golden_dir):
    """sample_optimize_density_CONLIN.cpp: 133 design iterations, objective 1.42395e-4, == Density_CONLIN.vtk."""
    g = np.load(os.path.join(golden_dir, "density_conlin.npz"))
    P = problems.cantilever2d(60, 40, opt_kind=problems.OPT_CONLIN)
    R = orc.simp_run(P.eq, P.coords, P.conn, P.fixed, P.loads, P.filter_kind, P.nbrs, P.opt_kind, P.optp(), P.params(),
                     500, np.full(P.nelem, 0.5))
    assert R["iters"] == 133
    assert abs(R["hist"][-1, 0] / P.scale0 - 1.42395e-4) < 1e-9
    assert np.abs(R["rho"] - g["rho"]).max() < 2e-6
    np.testing.assert_allclose(R["u"], g["u"], rtol=6e-6, atol=1e-12)
    np.testing.assert_allclose(R["r"], g["r"], rtol=6e-6, atol=1e-9)


def test_solid_hex8_vs_result_linear_vtk(golden_dir):
    g = np.load(os.path.join(golden_dir, "solid_linear.npz"))
    fixed = (g["fix_node"], g["fix_dof"], g["fix_val"])
    loads = (g["load_node"], g["load_dof"], g["load_val"])
    S, n2g, ufix, _ = orc.assemble(orc.EQ_SOLID, g["coords"], g["conn"], fixed, loads, np.full(len(g["conn"]), 210000.0), 0.3, 1.0)
    x, it, relres = S.solve(1, S.arrays()[3])
    assert relres < 1e-10
    u = np.where(n2g >= 0, x[np.maximum(n2g, 0)], ufix)
    assert abs(np.abs(u).max() - 1.48963) < 1e-5
    np.testing.assert_allclose(sig6(u), g["u"], rtol=2e-5, atol=1e-9)


@pytest.mark.skipif(not reflib.available(), reason="live reference not built (needs /root/reference)")
def test_port_vs_live_reference_fresh_inputs():
    rng = np.random.default_rng(7)
    reflib.set_num_threads(1)
    for eq, P in ((orc.EQ_HEAT, problems.heat2d(10, 6)), (orc.EQ_SOLID, problems.cantilever3d(4, 3, 2))):
        Emod = rng.uniform(0.5, 2.0, P.nelem)
        Sr = reflib.assemble(eq, P.coords, P.conn, P.fixed, P.loads, Emod)
        So, n2g, ufix, _ = orc.assemble(eq, P.coords, P.conn, P.fixed, P.loads, Emod)
        for a, b in zip(Sr.arrays(), So.arrays()):
            assert np.array_equal(a, b)
        b = So.arrays()[3]
        for kind in (0, 1, 2):
            xr = Sr.solve(kind, b)[0]
            xo = So.solve(kind, b)[0]
            np.testing.assert_allclose(xo, xr, rtol=0, atol=1e-13 * np.abs(xr).max())
    # 3-D SIMP loop, 3 iterations, MMA + density filter
    P = problems.cantilever3d(6, 4, 2, opt_kind=problems.OPT_MMA)
    a = reflib.simp_run(P.eq, P.coords, P.conn, P.fixed, P.loads, P.filter_kind, P.nbrs, P.opt_kind, P.optp(), P.params(), 3, np.full(P.nelem, 0.5), False)
    b = orc.simp_run(P.eq, P.coords, P.conn, P.fixed, P.loads, P.filter_kind, P.nbrs, P.opt_kind, P.optp(), P.params(), 3, np.full(P.nelem, 0.5), False)
    np.testing.assert_allclose(b["hist"][:, 0], a["hist"][:, 0], rtol=1e-9)
    assert np.abs(a["s"] - b["s"]).max() < 1e-8
