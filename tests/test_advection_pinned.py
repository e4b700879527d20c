"""Pin the C restatement of the advection-diffusion family (Advection.h:19-229) on live-reference
fixtures (tests/golden/live_advection.npz) and on the reference's committed outputs sample/advection/AdvectionSUPG.vtk,
AdvectionSUPGdynamic0.vtk and AdvectionSUPGdynamic99.vtk.  CPU only."""
import os

import numpy as np
import pytest

from oracle import portlib as orc
from oracle import reflib

T3, G1TRI = 1, 1
DYN_TERMS = 1 | 2 | 4 | 16 | 32      # sample_advectiondiffusion_dynamic.cpp:57-69: M, MS, A, D, AS


@pytest.fixture(scope="module")
def adv(golden_dir):
    return np.load(os.path.join(golden_dir, "live_advection.npz"))


def test_every_routine_on_every_shape_is_bit_identical(adv):
    cases = adv["cases"]
    assert len(cases) == 180
    for i, (shape, quad, terms, ax, ay, k) in enumerate(cases):
        ke = orc.advdiff_element(int(shape), int(quad), int(terms), adv[f"xe_{i}"], ax, ay, k)
        assert np.array_equal(ke, adv[f"ke_{i}"]), (i, shape, quad, terms)


def test_static_sample_system_and_vtk(adv, golden_dir):
    """sample_advectiondiffusion_static.cpp: Advection + Diffusion + AdvectionSUPG on the T3 mesh, a = 1 at 60 degrees, k = 1e-6."""
    kry = np.load(os.path.join(golden_dir, "live_krylov.npz"))
    ax, ay = 1.0 * np.cos(60.0 * np.pi / 180.0), 1.0 * np.sin(60.0 * np.pi / 180.0)
    vel = np.tile([ax, ay], (len(adv["smp_conn"]), 1))
    S, n2g, T = orc.advdiff_system(T3, G1TRI, 7, adv["smp_coords"], adv["smp_conn"], adv["smp_fix_node"], adv["smp_fix_val"], vel, 1.0e-6)
    indptr, indices, data, F = S.arrays()
    assert np.array_equal(indptr, kry["adv_indptr"]) and np.array_equal(indices, kry["adv_indices"])
    assert np.array_equal(data, kry["adv_data"]) and np.array_equal(F, kry["adv_F"])
    x, it, relres = S.solve(3, F)
    assert relres < 1e-10
    free = n2g[:, 0] >= 0
    T[free] = x[n2g[free, 0]]
    assert np.abs(T - adv["smp_T_static_vtk"]).max() < 5e-6 * max(1.0, np.abs(adv["smp_T_static_vtk"]).max())


def test_dynamic_sample_first_step_system_is_bit_identical(adv):
    S, n2g, T = orc.advdiff_system(T3, G1TRI, DYN_TERMS, adv["smp_coords"], adv["smp_conn"], adv["smp_fixd_node"], adv["smp_fixd_val"],
                                   adv["smp_vel"], 0.0, np.pi / 50.0, 0.5, adv["smp_T0"])
    indptr, indices, data, F = S.arrays()
    assert np.array_equal(n2g, adv["dyn_n2g"])
    assert np.array_equal(indptr, adv["dyn_indptr"]) and np.array_equal(indices, adv["dyn_indices"])
    assert np.array_equal(data, adv["dyn_data"]) and np.array_equal(F, adv["dyn_F"])


def test_dynamic_sample_100_steps_and_vtks(adv):
    """The whole run of sample_advectiondiffusion_dynamic.cpp (rotating cone, Crank-Nicolson, SUPG, BiCGSTAB) through the restatement."""
    T = adv["smp_T0"].copy()
    for step in range(100):
        S, n2g, T = orc.advdiff_system(T3, G1TRI, DYN_TERMS, adv["smp_coords"], adv["smp_conn"], adv["smp_fixd_node"], adv["smp_fixd_val"],
                                       adv["smp_vel"], 0.0, np.pi / 50.0, 0.5, T)
        x, it, relres = S.solve(3, S.arrays()[3])
        assert relres < 1e-10
        free = n2g[:, 0] >= 0
        T[free] = x[n2g[free, 0]]
        if step in (0, 1, 2, 99):
            assert np.abs(T - adv[f"dyn_T{step}"]).max() < 1e-11, step
    assert np.abs(T - adv["smp_T_dyn99_vtk"]).max() < 1e-6
    assert np.abs(adv["dyn_T0"] - adv["smp_T_dyn0_vtk"]).max() < 1e-6


def heat_dynamic_problem(adv):
    """sample/heattransfer/sample_heattransfer_dynamic.cpp:19-45: unit square (T3), T = 300 on x = 0, T = 0 on x = 1, conductivity = capacity = 1,
    dt = 0.001, theta = 0.5, 500 steps from T = 0.  HeatTransfer + HeatCapacity are the Diffusion + Mass pair of the advection family."""
    coords, conn = adv["heatdyn_coords"], adv["heatdyn_conn"]
    left = np.nonzero(np.abs(coords[:, 0]) < 1e-5)[0]
    right = np.nonzero(np.abs(coords[:, 0] - 1.0) < 1e-5)[0]
    fn = np.concatenate([left, right]).astype(np.int32)
    fv = np.concatenate([np.full(len(left), 300.0), np.zeros(len(right))])
    return coords, conn, fn, fv


def heat_dynamic_run(adv, steps=500):
    coords, conn, fn, fv = heat_dynamic_problem(adv)
    vel, T = np.zeros((len(conn), 2)), np.zeros(len(coords))
    for step in range(steps):
        S, n2g, T = orc.advdiff_system(T3, G1TRI, 2 | 16, coords, conn, fn, fv, vel, 1.0, 0.001, 0.5, T)
        x, it, relres = S.solve(1, S.arrays()[3])            # ScalingCG, as the sample
        assert relres < 1e-10
        free = n2g[:, 0] >= 0
        T[free] = x[n2g[free, 0]]
    return T


def test_heat_conduction_theta_scheme_reproduces_dynamic_vtk(adv):
    T = heat_dynamic_run(adv)
    # 6 printed digits of values up to 300, on coordinates that were themselves printed with 6 digits
    assert np.abs(T - adv["heatdyn_T_vtk"]).max() < 2e-3
    assert T.min() > -1e-9 and abs(T.max() - 300.0) < 1e-12


@pytest.mark.parametrize("nm", ["q4", "t6", "q8"])
def test_all_six_routines_time_step_on_family_meshes(adv, nm):
    S, n2g, T = orc.advdiff_system(int(adv[f"{nm}_shape"]), int(adv[f"{nm}_quad"]), 63, adv[f"{nm}_coords"], adv[f"{nm}_conn"], adv[f"{nm}_fix_node"],
                                   adv[f"{nm}_fix_val"], adv[f"{nm}_vel"], 0.02, 0.05, 0.6, adv[f"{nm}_Tn"])
    indptr, indices, data, F = S.arrays()
    assert np.array_equal(indptr, adv[f"{nm}_indptr"]) and np.array_equal(indices, adv[f"{nm}_indices"])
    assert np.array_equal(data, adv[f"{nm}_data"]) and np.array_equal(F, adv[f"{nm}_F"])
    x, it, relres = S.solve(3, F)
    np.testing.assert_allclose(x, adv[f"{nm}_x"], rtol=0, atol=1e-11 * np.abs(adv[f"{nm}_x"]).max())


@pytest.mark.skipif(not reflib.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_against_the_live_reference_on_fresh_inputs(adv):
    rng = np.random.default_rng(99)
    for i in rng.choice(180, 24, replace=False):
        shape, quad, terms, ax, ay, k = adv["cases"][i]
        xe = adv[f"xe_{i}"] + 0.03 * rng.uniform(-1, 1, adv[f"xe_{i}"].shape)
        a = orc.advdiff_element(int(shape), int(quad), int(terms), xe, ax + 0.1, ay - 0.2, k)
        b = reflib.advdiff_element(int(shape), int(quad), int(terms), xe, ax + 0.1, ay - 0.2, k)
        assert np.array_equal(a, b)
