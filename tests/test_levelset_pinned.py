"""Pin the C restatement of the level-set loop (oracle/pf2_oracle.c orc_levelset_run) against
  * the UNMODIFIED sample/optimize/sample_optimize_levelset.cpp: its printed history (6 digits), the iteration at which its
    convergence test fires (result files 0..115) and the fields of its last VTK, and
  * the same loop run through the reference's own routines at full precision (oracle/ref_shim.cpp ref_levelset_run).
Fixtures: tests/golden/levelset.npz (tests/golden/make_golden.py levelset).  CPU only."""
import os

import numpy as np
import pytest

from oracle import portlib as orc
from pansfem2_b200 import problems


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "levelset.npz"))


@pytest.fixture(scope="module")
def full_run():
    P = problems.levelset2d()
    return P, orc.levelset_run(P.coords, P.conn, P.fixed, P.loads, P.phifixed, P.prm(), P.tmax, np.ones(P.nnode), np.ones(P.nelem))


def test_full_run_vs_unmodified_sample(gold, full_run):
    P, R = full_run
    so = gold["stdout_hist"]
    assert R["converged"] and R["iters"] == 116 == int(gold["vtk_count"]) and len(so) == 115       # t = 0..114 printed, test fires at 115
    mine = np.stack([R["hist"][:115, 0] / P.nelem, R["hist"][:115, 1], R["hist"][:115, 2]], axis=1)
    np.testing.assert_allclose(mine, so[:, 1:], rtol=6e-6)                                         # 6 printed digits
    assert np.array_equal(R["str"], gold["vtk_str"])
    np.testing.assert_allclose(R["phi"], gold["vtk_phi"], rtol=6e-6, atol=1e-12)
    np.testing.assert_allclose(R["u"], gold["vtk_u"], rtol=6e-6, atol=1e-9)


def test_full_run_vs_reference_routines_full_precision(gold, full_run):
    P, R = full_run
    assert R["iters"] == int(gold["iters"])
    np.testing.assert_allclose(R["hist"], gold["hist"], rtol=1e-12)
    assert np.array_equal(R["str"], gold["str"])
    assert np.abs(R["phi"] - gold["phi"]).max() < 1e-12
    np.testing.assert_allclose(R["u"], gold["u"], rtol=0, atol=1e-12 * np.abs(gold["u"]).max())


def test_small_case_from_perturbed_state(gold):
    P = problems.levelset2d(24, 16, nvol=10.0)
    R = orc.levelset_run(P.coords, P.conn, P.fixed, P.loads, P.phifixed, P.prm(), 25, gold["small_phi0"], gold["small_str0"])
    assert R["iters"] == int(gold["small_iters"])
    np.testing.assert_allclose(R["hist"], gold["small_hist"], rtol=1e-12)
    assert np.array_equal(R["str"], gold["small_str"]) and np.abs(R["phi"] - gold["small_phi"]).max() < 1e-12
