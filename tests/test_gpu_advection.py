"""GPU parity of the advection-diffusion family (Advection.h:19-229, sample/advection) through the C ABI
and the C++ header mirror: every routine on every 2-D <SF, IC> against live-reference fixtures (tests/golden/live_advection.npz),
pf2_advdiff_assemble against the reference-assembled systems of both advection samples and of Q4 / T6 / Q8 time steps, the device
time loop against the live-reference history, and the committed AdvectionSUPG.vtk / AdvectionSUPGdynamic{0,99}.vtk."""
import os
import subprocess

import numpy as np
import pytest

from pansfem2_b200 import capi
from pansfem2_b200 import eqcode as ec


pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "pansfem2_b200", "bin")
DYN_TERMS = ec.ADV_ADVECTION | ec.ADV_DIFFUSION | ec.ADV_SUPG | ec.ADV_MASS | ec.ADV_MASS_SUPG


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def adv(golden_dir):
    return np.load(os.path.join(golden_dir, "live_advection.npz"))


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


class Problem:
    def __init__(self, ctx, coords, conn, fix_node, fix_val):
        self.ctx = ctx
        self.mesh = capi.Mesh(ctx, coords, conn)
        fn = np.asarray(fix_node, np.int32)
        self.map = capi.DofMap(ctx, len(coords), 1, (fn, np.zeros_like(fn), np.asarray(fix_val, float)))
        self.K = capi.Csr.pattern(ctx, self.mesh, self.map)

    def close(self):
        self.K.close(); self.map.close(); self.mesh.close()


def test_every_routine_on_every_shape(ctx, adv):
    cases = adv["cases"]
    assert len(cases) == 180
    for i, (shape, quad, terms, ax, ay, k) in enumerate(cases):
        eq = ec.eq_code(ec.PHYS_ADVDIFF, int(shape), int(quad), int(terms))
        ke = ctx.element_matrix(eq, adv[f"xe_{i}"], ax, ay, k)
        assert rel(ke, adv[f"ke_{i}"]) < 1e-13, ec.describe(eq)


def test_invalid_selections_are_rejected(ctx):
    xe = np.zeros((8, 2))
    for eq in (ec.eq_code(ec.PHYS_ADVDIFF, ec.SHAPE_T3, ec.QUAD_G1TRI, 0), ec.eq_code(ec.PHYS_ADVDIFF, ec.SHAPE_T3, ec.QUAD_G4SQ, 1),
               ec.eq_code(ec.PHYS_ADVDIFF, ec.SHAPE_TET4, ec.QUAD_G1TET, 1), ec.eq_code(ec.PHYS_ADVDIFF, ec.SHAPE_Q4, ec.QUAD_G4SQ, 64)):
        with pytest.raises(capi.Pf2Error):
            ctx.element_matrix(eq, xe, 1.0)


def test_static_sample_assembled_and_solved_on_the_device(ctx, adv, golden_dir):
    """sample_advectiondiffusion_static.cpp:37-58 as pf2_advdiff_assemble + BiCGSTAB + pf2_disassemble -> AdvectionSUPG.vtk."""
    kry = np.load(os.path.join(golden_dir, "live_krylov.npz"))
    P = Problem(ctx, adv["smp_coords"], adv["smp_conn"], adv["smp_fix_node"], adv["smp_fix_val"])
    eq = ec.eq_code(ec.PHYS_ADVDIFF, ec.SHAPE_T3, ec.QUAD_G1TRI, 7)
    ax, ay = 1.0 * np.cos(60.0 * np.pi / 180.0), 1.0 * np.sin(60.0 * np.pi / 180.0)
    P.K.advdiff_assemble(P.mesh, P.map, eq, (ax, ay, 1.0e-6, 0.0, 1.0, 0.0))
    indptr, indices, data, F = P.K.download()
    assert np.array_equal(indptr, kry["adv_indptr"]) and np.array_equal(indices, kry["adv_indices"])
    assert rel(data, kry["adv_data"]) < 1e-13 and rel(F, kry["adv_F"]) < 1e-13
    x, T = ctx.empty(P.K.rows), ctx.empty(len(adv["smp_coords"]))
    for solver in (capi.SOLVER_BICGSTAB, capi.SOLVER_SCALINGBICGSTAB):
        it, relres = P.K.solve(solver, P.K.device_F(), x)
        assert relres < 1e-10
        P.map.disassemble(x, T)
        np.testing.assert_allclose(T.download(), adv["smp_T_static_vtk"], rtol=6e-6, atol=2e-6)
    P.close()


def test_dynamic_sample_time_loop_on_the_device(ctx, adv):
    """sample_advectiondiffusion_dynamic.cpp:44-75: 100 Crank-Nicolson SUPG steps of the rotating cone, field resident on the device."""
    P = Problem(ctx, adv["smp_coords"], adv["smp_conn"], adv["smp_fixd_node"], adv["smp_fixd_val"])
    assert np.array_equal(P.map.get(), adv["dyn_n2g"])
    eq = ec.eq_code(ec.PHYS_ADVDIFF, ec.SHAPE_T3, ec.QUAD_G1TRI, DYN_TERMS)
    dt, theta = np.pi / 50.0, 0.5
    prm = (0.0, 0.0, 0.0, 1.0 / dt, theta, 1.0 - theta)
    vel, T, x = ctx.array(adv["smp_vel"].ravel()), ctx.array(adv["smp_T0"]), ctx.empty(P.K.rows)
    launches0 = ctx.launch_count()
    for step in range(100):
        P.K.advdiff_assemble(P.mesh, P.map, eq, prm, vel=vel, T=T)
        if step == 0:
            indptr, indices, data, F = P.K.download()
            assert np.array_equal(indptr, adv["dyn_indptr"]) and np.array_equal(indices, adv["dyn_indices"])
            assert rel(data, adv["dyn_data"]) < 1e-13 and rel(F, adv["dyn_F"]) < 1e-13
        it, relres = P.K.solve(capi.SOLVER_BICGSTAB, P.K.device_F(), x)
        assert relres < 1e-10
        P.map.disassemble(x, T)
        if step in (0, 1, 2, 99):
            assert np.abs(T.download() - adv[f"dyn_T{step}"]).max() < 1e-8, step
    assert ctx.launch_count() > launches0 + 300
    Tn = T.download()
    assert np.abs(Tn - adv["smp_T_dyn99_vtk"]).max() < 1e-6
    P.close()


def test_heat_conduction_theta_scheme_on_the_device(ctx, adv):
    """sample/heattransfer/sample_heattransfer_dynamic.cpp (HeatTransfer + HeatCapacity, Crank-Nicolson, ScalingCG, 500 steps) through the
    Diffusion + Mass pair of pf2_advdiff_assemble with the field resident on the device -> the committed dynamic.vtk and the oracle's run."""
    from test_advection_pinned import heat_dynamic_problem, heat_dynamic_run
    coords, conn, fn, fv = heat_dynamic_problem(adv)
    P = Problem(ctx, coords, conn, fn, fv)
    eq = ec.eq_code(ec.PHYS_ADVDIFF, ec.SHAPE_T3, ec.QUAD_G1TRI, ec.ADV_DIFFUSION | ec.ADV_MASS)
    dt, theta = 0.001, 0.5
    prm = (0.0, 0.0, 1.0, 1.0 / dt, theta, 1.0 - theta)
    T, x = ctx.array(np.zeros(len(coords))), ctx.empty(P.K.rows)
    for step in range(500):
        P.K.advdiff_assemble(P.mesh, P.map, eq, prm, T=T)
        it, relres = P.K.solve(capi.SOLVER_SCALINGCG, P.K.device_F(), x)
        assert relres < 1e-10
        P.map.disassemble(x, T)
    Tn = T.download()
    assert np.abs(Tn - adv["heatdyn_T_vtk"]).max() < 2e-3
    assert np.abs(Tn - heat_dynamic_run(adv)).max() < 1e-4
    P.close()


@pytest.mark.parametrize("nm", ["q4", "t6", "q8"])
def test_all_six_routines_time_step_on_family_meshes(ctx, adv, nm):
    P = Problem(ctx, adv[f"{nm}_coords"], adv[f"{nm}_conn"], adv[f"{nm}_fix_node"], adv[f"{nm}_fix_val"])
    eq = ec.eq_code(ec.PHYS_ADVDIFF, int(adv[f"{nm}_shape"]), int(adv[f"{nm}_quad"]), 63)
    k, dt, theta = 0.02, 0.05, 0.6
    vel, T = ctx.array(adv[f"{nm}_vel"].ravel()), ctx.array(adv[f"{nm}_Tn"])
    P.K.advdiff_assemble(P.mesh, P.map, eq, (0.0, 0.0, k, 1.0 / dt, theta, 1.0 - theta), vel=vel, T=T)
    indptr, indices, data, F = P.K.download()
    assert np.array_equal(indptr, adv[f"{nm}_indptr"]) and np.array_equal(indices, adv[f"{nm}_indices"])
    assert rel(data, adv[f"{nm}_data"]) < 1e-13 and rel(F, adv[f"{nm}_F"]) < 1e-12
    x, it, relres = P.K.solve_host(capi.SOLVER_BICGSTAB, F)
    assert relres < 1e-10
    assert np.abs(x - adv[f"{nm}_x"]).max() < 1e-8 * np.abs(adv[f"{nm}_x"]).max()
    # pf2_assemble / pf2_compliance_sens refuse the non-symmetric selections
    with pytest.raises(capi.Pf2Error):
        P.K.assemble(P.mesh, P.map, eq, (1.0, 1.0, 0.3, 1.0, 1.0), (np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0)), modulus=T)
    P.close()


# ---- the C++ side: batched driver and the reference's UNMODIFIED samples on the header mirror ----------------------------------
def parse_T(path, n):
    L = open(path).read().split("\n")
    i0 = next(k for k, ln in enumerate(L) if ln.startswith("SCALARS T"))
    return np.array([float(v) for v in L[i0 + 2:i0 + 2 + n]])


def write_model(d, adv):
    """sample/advection/{Node,Element,Dirichlet,DirichletD}.csv from the fixture (the reference tree is not on the GPU box)."""
    d.mkdir(parents=True, exist_ok=True)
    with open(d / "Node.csv", "w") as f:
        f.write("ID,x0,x1\n")
        for i, c in enumerate(adv["smp_coords"]):
            f.write(f"{i},{float(c[0])!r},{float(c[1])!r}\n")
    with open(d / "Element.csv", "w") as f:
        f.write("Cell ID,Point Index 0,Point Index 1,Point Index 2\n")
        for i, e in enumerate(adv["smp_conn"]):
            f.write(f"{i},{e[0]},{e[1]},{e[2]}\n")
    for name, nn, vv in (("Dirichlet.csv", adv["smp_fix_node"], adv["smp_fix_val"]), ("DirichletD.csv", adv["smp_fixd_node"], adv["smp_fixd_val"])):
        with open(d / name, "w") as f:
            f.write("ID,T\n")
            for n, v in zip(nn, vv):
                f.write(f"{int(n)},{float(v)!r}\n")


def need(name):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (python -m pansfem2_b200.build_cpp)")
    return exe


@pytest.mark.parametrize("mode,golden", [("static", "smp_T_static_vtk"), ("dynamic", "smp_T_dyn99_vtk")])
def test_batched_cpp_driver(tmp_path, adv, mode, golden):
    exe = need("sample_advectiondiffusion_batched")
    write_model(tmp_path / "model", adv)
    out = tmp_path / "result.vtk"
    r = subprocess.run([exe, mode, str(tmp_path / "model"), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    T = parse_T(out, len(adv["smp_coords"]))
    np.testing.assert_allclose(T, adv[golden], rtol=6e-6, atol=2e-6)


def test_unmodified_reference_static_sample_on_the_header_mirror(tmp_path, adv):
    exe = need("dropin_advection_static")
    write_model(tmp_path / "sample" / "advection", adv)
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    T = parse_T(tmp_path / "sample" / "advection" / "result.vtk", len(adv["smp_coords"]))
    np.testing.assert_allclose(T, adv["smp_T_static_vtk"], rtol=6e-6, atol=2e-6)


def test_unmodified_reference_dynamic_sample_on_the_header_mirror(tmp_path, adv):
    """100 steps x 800 elements x 5 per-element device calls: the legacy per-element path, kept for compatibility."""
    exe = need("dropin_advection_dynamic")
    write_model(tmp_path / "sample" / "advection", adv)
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    n = len(adv["smp_coords"])
    np.testing.assert_allclose(parse_T(tmp_path / "sample" / "advection" / "result0.vtk", n), adv["smp_T_dyn0_vtk"], rtol=6e-6, atol=2e-6)
    np.testing.assert_allclose(parse_T(tmp_path / "sample" / "advection" / "result99.vtk", n), adv["smp_T_dyn99_vtk"], rtol=6e-6, atol=2e-6)
