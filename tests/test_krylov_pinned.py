"""Pin the C restatement of the non-symmetric Krylov family (orc_solve_bicgstab: BiCGSTAB, BiCGSTAB2, ScalingBiCGSTAB,
ILU0BiCGSTAB) against solutions of the live reference (tests/golden/live_krylov.npz).  CPU only."""
import os

import numpy as np
import pytest

from oracle import portlib as orc

KINDS = [(3, "bicgstab"), (4, "bicgstab2"), (5, "scalingbicgstab"), (6, "ilu0bicgstab")]


@pytest.fixture(scope="module")
def kry(golden_dir):
    return np.load(os.path.join(golden_dir, "live_krylov.npz"))


@pytest.mark.parametrize("kind,nm", KINDS)
def test_port_vs_live_reference_fixture(kry, kind, nm):
    A = orc.system_from_csr(kry["indptr"], kry["indices"], kry["data"])
    x, it, relres = A.solve(kind, kry["b"])
    assert relres < 1e-10 and 5 < it < 100
    assert np.array_equal(x, kry[f"x_{nm}"])                 # same recurrences, same operand order: bit-identical
    assert np.abs(x - kry["x_exact"]).max() < 1e-8 * np.abs(kry["x_exact"]).max()


def test_nonconvergence_returns_last_iterate(kry):
    A = orc.system_from_csr(kry["indptr"], kry["indices"], kry["data"])
    x, it, relres = A.solve(3, kry["b"], itrmax=5)
    assert it == 5 and relres > 1e-10 and np.isfinite(x).all()


def test_advection_supg_sample_vtk(kry):
    """sample/advection/sample_advectiondiffusion_static.cpp: K, F assembled by the reference's Advection.h routines (fixture), solved
    with BiCGSTAB as the sample does -> the committed AdvectionSUPG.vtk at its 6 printed digits."""
    A = orc.system_from_csr(kry["adv_indptr"], kry["adv_indices"], kry["adv_data"])
    x, it, relres = A.solve(3, kry["adv_F"])
    assert relres < 1e-10 and np.array_equal(x, kry["adv_x_bicgstab"])
    n2g = kry["adv_n2g"][:, 0]
    fixed = np.zeros(len(n2g))
    fixed[kry["adv_fix_node"]] = kry["adv_fix_val"]
    T = np.where(n2g >= 0, x[np.maximum(n2g, 0)], fixed)
    np.testing.assert_allclose(T, kry["adv_T"], rtol=6e-6, atol=1e-6)
