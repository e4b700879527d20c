"""The periodic cell problem of sample/homogenization/sample_homogenization.cpp (SquareAnnulusMesh2 20 x 20 with a 10 x 10 hole, periodic
pairs by SetPeriodic, PlaneStrainStiffness<Q4, Gauss4> + HomogenizePlaneStrainBodyForce + WeakSpring, three ScalingCG solves,
HomogenizePlaneStrainConstitutive) replayed with the restatement's element matrices and ScalingCG and numpy for the host-side integrals:
it must reproduce the committed result_microscopic.vtk and the matrices the unmodified sample prints (tests/golden/homogenization.npz).
The assembly order differs from the reference's LILCSR order and the rigid-body mode is only held by the 1e-9 weak spring, so this is
also the tolerance study for the device run of the unmodified driver (tests/test_gpu_cpp_dropin.py).  CPU only."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import portlib as orc
from pansfem2_b200 import eqcode as ec

E, V = 100.0, 0.3          # sample_homogenization.cpp:23-25


@pytest.fixture(scope="module")
def hom(golden_dir):
    return np.load(os.path.join(golden_dir, "homogenization.npz"))


def q4_strain_matrices(xe):
    """B (3 x 8) and det J at the four Gauss points (ShapeFunction4Square / Gauss4Square)."""
    a = 1.0 / np.sqrt(3.0)
    for r0, r1 in ((-a, -a), (a, -a), (-a, a), (a, a)):
        dN = np.array([[-0.25 * (1 - r1), 0.25 * (1 - r1), 0.25 * (1 + r1), -0.25 * (1 + r1)],
                       [-0.25 * (1 - r0), -0.25 * (1 + r0), 0.25 * (1 + r0), 0.25 * (1 - r0)]])
        J = dN @ xe
        g = np.linalg.inv(J) @ dN
        B = np.zeros((3, 8))
        B[0, 0::2], B[1, 1::2], B[2, 0::2], B[2, 1::2] = g[0], g[1], g[1], g[0]
        yield B, np.linalg.det(J)


def test_cell_problem_reproduces_result_microscopic_vtk(hom):
    coords, conn, pairs = hom["coords"], hom["conn"], hom["pairs"]
    n = len(coords)
    # SetPeriodic (BoundaryCondition.h:38-62)
    n2g = np.zeros((n, 2), int)
    n2g[pairs[:, 1]] = -1
    free = n2g.ravel() != -1
    n2g.ravel()[free] = np.arange(free.sum())
    n2g[pairs[:, 1]] = n2g[pairs[:, 0]]
    k = int(free.sum())
    D = np.array([[1 - V, V, 0], [V, 1 - V, 0], [0, 0, 0.5 * (1 - 2 * V)]]) * E / ((1 - 2 * V) * (1 + V))
    rows, cols, vals, F = [], [], [], np.zeros((3, k))
    eq = ec.eq_code(ec.PHYS_PLANESTRAIN)
    for el in conn:
        xe, dofs = coords[el], n2g[el].ravel()
        Ke = orc.element_matrix(eq, xe, E, V, 1.0)
        rows += list(np.repeat(dofs, 8)); cols += list(np.tile(dofs, 8)); vals += list(Ke.ravel())
        Fes = sum(B.T @ D * J for B, J in q4_strain_matrices(xe))                 # Homogenization.h:42-56
        for c in range(3):
            np.add.at(F[c], dofs, Fes[:, c])
        # WeakSpring (General.h:123-134): alpha I, and BOTH dofs of node i mapped to local column 2 i
        for i in range(4):
            for di in range(2):
                for dj in range(2):
                    rows.append(n2g[el[i], di]); cols.append(n2g[el[i], dj]); vals.append(1.0e-9)
    K = sp.coo_matrix((vals, (rows, cols)), shape=(k, k)).tocsr()
    K.sort_indices()
    S = orc.system_from_csr(K.indptr, K.indices, K.data)
    chi = []
    for c in range(3):
        x, it, relres = S.solve(1, F[c])                                           # ScalingCG
        assert relres < 1e-10 and it < 200
        chi.append(x[n2g])
        assert np.abs(chi[c] - hom[f"chi{c}"]).max() < 2e-6                        # 6 printed digits of values up to 0.28
        assert np.abs(chi[c].mean(axis=0)).max() < 1e-12                           # the rigid-body mode is not excited
    CH, check = np.zeros((3, 3)), np.eye(3)
    for el in conn:
        CHI = np.stack([chi[c][el].ravel() for c in range(3)], axis=1)            # 8 x 3
        for B, J in q4_strain_matrices(coords[el]):
            CH += D @ (np.eye(3) - B @ CHI) * J                                    # Homogenization.h:84-98
            check += -B @ CHI * J                                                  # :120-134 (the sample starts from the identity)
    np.testing.assert_allclose(CH, hom["CH"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(check, hom["check"], rtol=2e-5, atol=1e-6)
