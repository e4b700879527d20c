"""GPU parity of the level-set design loop (pf2_levelset_*) through the C ABI, against the fixtures of
tests/golden/levelset.npz: the reference's routines at full precision and the unmodified sample's own output."""
import os

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
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "levelset.npz"))


def test_small_case_from_perturbed_state(ctx, gold):
    P = problems.levelset2d(24, 16, nvol=10.0, tmax=25)
    L = capi.LevelSet(ctx, P)
    L.set_state(gold["small_phi0"], gold["small_str0"])
    hist = []
    for t in range(25):
        st = L.iterate()
        assert st["cg_relres"] < 1e-10 and not st["converged"]
        hist.append((st["objective"], st["vol"], st["lam"]))
    np.testing.assert_allclose(np.array(hist), gold["small_hist"], rtol=1e-8)
    out = L.get()
    assert np.array_equal(out["str"], gold["small_str"])
    assert np.abs(out["phi"] - gold["small_phi"]).max() < 1e-6
    L.close()


@pytest.mark.parametrize("matrix_free", [False, True])
def test_sample_run_to_convergence(ctx, gold, matrix_free):
    """sample_optimize_levelset.cpp on the device: 60 x 40, converges at t = 115 like the reference; history against the
    full-precision reference run, final fields against the sample's last VTK.  Also with K applied matrix-free."""
    P = problems.levelset2d()
    L = capi.LevelSet(ctx, P, matrix_free=matrix_free)
    hist = []
    t = 0
    for t in range(P.tmax):
        st = L.iterate()
        hist.append((st["objective"], st["vol"], st["lam"]))
        if t == 39:
            out = L.get()
            assert np.array_equal(out["str"], gold["str40"]) and np.abs(out["phi"] - gold["phi40"]).max() < 1e-6
        if st["converged"]:
            break
    hist = np.array(hist)
    assert t + 1 == int(gold["iters"]) == 116
    np.testing.assert_allclose(hist, gold["hist"], rtol=1e-8)
    out = L.get()
    assert np.array_equal(out["str"], gold["vtk_str"])
    np.testing.assert_allclose(out["phi"], gold["vtk_phi"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(out["u"], gold["vtk_u"], rtol=1e-5, atol=1e-9)
    L.close()
