"""Host-side structured meshers and filter neighbour lists (numpy; set-up only, off the timed path).

Numbering follows the reference so that every fixture, golden file and partition lines up:

* ``square_mesh``  mirrors ``SquareMesh<T>::GenerateNodes/GenerateElements``
  (/root/reference/src/PrePost/Mesher/SquareMesh.h:62-107): node id ``(ny+1)*i + j``, element id
  ``ny*i + j``, nodes counter-clockwise ``(i,j) (i+1,j) (i+1,j+1) (i,j+1)``.
* ``box_mesh`` is OUR x-major hex8 mesher (the reference has no 3-D mesher):
  node id ``((ny+1)*i + j)*(nz+1) + k``, element id ``(ny*i + j)*nz + k``, node order = bottom face CCW then
  top face CCW as ``ShapeFunction8Cubic`` expects (ShapeFunction.h:299).
* ``filter_neighbors_*`` reproduce what the samples build with an all-pairs search
  (sample/optimize/sample_optimize_density_oc.cpp:50-60): neighbours with centroid distance ``d <= R``
  in ascending element id, weight ``(R-d)/R``; built structurally so it scales to millions of elements.
"""
from __future__ import annotations

import numpy as np


def square_mesh(lx: float, ly: float, nx: int, ny: int, xr=None):
    """xr = (i0, i1): only element columns [i0, i1) (node columns [i0, i1]) with LOCAL ids - one slab of the mesh."""
    i0, i1 = xr if xr is not None else (0, nx)
    i, j = np.meshgrid(np.arange(i0, i1 + 1), np.arange(ny + 1), indexing="ij")
    coords = np.empty(((i1 - i0 + 1) * (ny + 1), 2), dtype=np.float64)
    coords[:, 0] = (lx * (i / float(nx))).ravel()
    coords[:, 1] = (ly * (j / float(ny))).ravel()
    ei, ej = np.meshgrid(np.arange(i1 - i0), np.arange(ny), indexing="ij")
    n0 = ((ny + 1) * ei + ej).ravel()
    conn = np.stack([n0, n0 + (ny + 1), n0 + (ny + 1) + 1, n0 + 1], axis=1).astype(np.int32)
    return coords, conn


def box_mesh(lx: float, ly: float, lz: float, nx: int, ny: int, nz: int, xr=None):
    i0, i1 = xr if xr is not None else (0, nx)
    i, j, k = np.meshgrid(np.arange(i0, i1 + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    coords = np.empty(((i1 - i0 + 1) * (ny + 1) * (nz + 1), 3), dtype=np.float64)
    coords[:, 0] = (lx * (i / float(nx))).ravel()
    coords[:, 1] = (ly * (j / float(ny))).ravel()
    coords[:, 2] = (lz * (k / float(nz))).ravel()
    ei, ej, ek = np.meshgrid(np.arange(i1 - i0), np.arange(ny), np.arange(nz), indexing="ij")
    sx, sy = (ny + 1) * (nz + 1), (nz + 1)
    n0 = (ei * sx + ej * sy + ek).ravel().astype(np.int64)
    conn = np.stack([n0, n0 + sx, n0 + sx + sy, n0 + sy,
                     n0 + 1, n0 + sx + 1, n0 + sx + sy + 1, n0 + sy + 1], axis=1).astype(np.int32)
    return coords, conn


def fixed_list(coords: np.ndarray, dofs, predicate, value: float = 0.0):
    """Mirror of ``SquareMesh<T>::GenerateFixedlist`` (SquareMesh.h:194-207): node-major, dof-minor list of
    ((node, dof), value) for every node whose coordinates satisfy ``predicate``."""
    mask = predicate(coords)
    nodes = np.nonzero(mask)[0].astype(np.int32)
    dofs = np.asarray(dofs, dtype=np.int32)
    node = np.repeat(nodes, len(dofs))
    dof = np.tile(dofs, len(nodes))
    val = np.full(node.shape, value, dtype=np.float64)
    return node, dof, val


def _stencil_neighbors(shape, radius: float, h):
    """Ragged (CSR) neighbour lists for a structured grid of cells with spacing ``h`` per axis."""
    dim = len(shape)
    reach = [int(np.floor(radius / h[d] + 1e-12)) for d in range(dim)]
    offs = np.stack(np.meshgrid(*[np.arange(-r, r + 1) for r in reach], indexing="ij"), axis=-1).reshape(-1, dim)
    dist = np.sqrt(((offs * np.asarray(h)) ** 2).sum(axis=1))
    keep = dist <= radius
    offs, dist = offs[keep], dist[keep]
    strides = np.ones(dim, dtype=np.int64)
    for d in range(dim - 2, -1, -1):
        strides[d] = strides[d + 1] * shape[d + 1]
    # ascending element id == lexicographic offsets (x-major), which is how meshgrid(indexing="ij") orders them
    order = np.argsort(offs @ strides, kind="stable")
    offs, dist = offs[order], dist[order]
    n = int(np.prod(shape))
    idx = np.stack(np.unravel_index(np.arange(n, dtype=np.int64), shape), axis=1)      # n x dim
    nb = idx[:, None, :] + offs[None, :, :]                                           # n x k x dim
    valid = np.ones(nb.shape[:2], dtype=bool)
    for d in range(dim):
        valid &= (nb[:, :, d] >= 0) & (nb[:, :, d] < shape[d])
    ids = (nb * strides).sum(axis=2)
    wts = np.broadcast_to(((radius - dist) / radius)[None, :], ids.shape)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(valid.sum(axis=1))
    return rowptr, ids[valid].astype(np.int32), np.ascontiguousarray(wts[valid], dtype=np.float64)


def filter_neighbors_2d(nx: int, ny: int, lx: float, ly: float, radius: float):
    return _stencil_neighbors((nx, ny), radius, (lx / nx, ly / ny))


def filter_neighbors_3d(nx: int, ny: int, nz: int, lx: float, ly: float, lz: float, radius: float):
    return _stencil_neighbors((nx, ny, nz), radius, (lx / nx, ly / ny, lz / nz))


def filter_neighbors_allpairs(centroids: np.ndarray, radius: float):
    """The sample's O(n^2) search verbatim (small meshes only) - used to validate the structural builders."""
    n = len(centroids)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    nbr, w = [], []
    for i in range(n):
        d = np.sqrt(((centroids - centroids[i]) ** 2).sum(axis=1))
        js = np.nonzero(d <= radius)[0]
        nbr.append(js)
        w.append((radius - d[js]) / radius)
        rowptr[i + 1] = rowptr[i] + len(js)
    return rowptr, np.concatenate(nbr).astype(np.int32), np.concatenate(w).astype(np.float64)


# ------------------------------------------------------------------------------------------------------------------
# Element families beyond Q4 / hex8.  Node order inside an element follows the reference's
# ShapeFunction*::Points (ShapeFunction.h:98, 133, 171, 207, 262, 299, 345-365).
# ------------------------------------------------------------------------------------------------------------------
NATURAL_NODES = {
    "T3": np.array([[1, 0], [0, 1], [0, 0]], float),
    "T6": np.array([[1, 0], [0, 1], [0, 0], [0.5, 0.5], [0, 0.5], [0.5, 0]], float),
    "Q4": np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], float),
    "Q8": np.array([[-1, -1], [1, -1], [1, 1], [-1, 1], [0, -1], [1, 0], [0, 1], [-1, 0]], float),
    "Tet4": np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, 0]], float),
    "Hex8": np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], float),
    # the 20 nodes ShapeFunction20Cubic::dNdr differentiates (ShapeFunction.h:396-461): corners, then mid-edges
    # 8..11 bottom face, 12..15 top face, 16..19 vertical edges
    "Hex20": np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1],
                       [0, -1, -1], [1, 0, -1], [0, 1, -1], [-1, 0, -1], [0, -1, 1], [1, 0, 1], [0, 1, 1], [-1, 0, 1],
                       [-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], float),
}


def _weld(points: np.ndarray, cells: int, npe: int):
    """Merge coincident element nodes (coordinates are multiples of 1/2 cell): returns (coords, conn)."""
    key = np.round(points * 4.0).astype(np.int64)
    uniq, inv = np.unique(key, axis=0, return_inverse=True)
    return uniq.astype(np.float64) / 4.0, inv.reshape(cells, npe).astype(np.int32)


def family_mesh(family: str, n, lengths=None):
    """Structured mesh of an n[0] x n[1] (x n[2]) block of unit-aspect cells made of `family` elements:
    T3 / T6 = each cell cut in two along its diagonal, Q8, Tet4 = each cell cut in six (Kuhn), Hex20.
    Returns (coords, conn); lengths default to the cell counts (unit cells)."""
    n = tuple(int(v) for v in n)
    dim = len(n)
    lengths = tuple(float(v) for v in (lengths or n))
    nat = NATURAL_NODES[family]
    cells = np.stack(np.meshgrid(*[np.arange(v) for v in n], indexing="ij"), axis=-1).reshape(-1, dim).astype(float)
    if family in ("Q4", "Q8", "Hex8", "Hex20"):
        local = [(nat + 1.0) / 2.0]                                    # one element per cell on [0,1]^dim
    elif family in ("T3", "T6"):
        # two counter-clockwise triangles; vertex i of the triangle sits at natural point i (r0, r1, 1-r0-r1)
        tris = [np.array([[1, 0], [1, 1], [0, 0]], float), np.array([[1, 1], [0, 1], [0, 0]], float)]
        local = [np.array([p[0] * v[0] + p[1] * v[1] + (1 - p[0] - p[1]) * v[2] for p in nat]) for v in tris]
    elif family == "Tet4":
        import itertools
        local = []
        for perm in itertools.permutations(range(3)):                  # Kuhn subdivision: 6 tetrahedra per cube
            v = [np.zeros(3)]
            for ax in perm:
                w = v[-1].copy(); w[ax] = 1.0; v.append(w)
            t = np.array([v[1], v[2], v[3], v[0]])
            if np.linalg.det(t[:3] - t[3]) < 0:
                t = t[[1, 0, 2, 3]]
            local.append(t)
    else:
        raise ValueError(family)
    pts = np.concatenate([(cells[:, None, :] + l[None, :, :]) for l in local], axis=1)     # cells x (k*npe) x dim
    npe = nat.shape[0]
    coords, conn = _weld(pts.reshape(-1, dim), cells.shape[0] * len(local), npe)
    coords = coords * (np.array(lengths) / np.array(n, float))
    return coords, conn


def element_centroids(coords: np.ndarray, conn: np.ndarray):
    """CenterOfGravity (General.h): mean of the element's node coordinates."""
    return coords[conn].mean(axis=1)


def filter_neighbors_centroid(centroids: np.ndarray, radius: float):
    """Same lists as the all-pairs search of the samples (ascending element id, weight (R-d)/R) through a k-d tree."""
    from scipy.spatial import cKDTree
    tree = cKDTree(centroids)
    lists = tree.query_ball_point(centroids, radius * (1.0 + 1e-12))
    rowptr = np.zeros(len(centroids) + 1, dtype=np.int64)
    nbr, w = [], []
    for i, js in enumerate(lists):
        js = np.array(sorted(js), dtype=np.int64)
        d = np.sqrt(((centroids[js] - centroids[i]) ** 2).sum(axis=1))
        keep = d <= radius
        js, d = js[keep], d[keep]
        nbr.append(js); w.append((radius - d) / radius)
        rowptr[i + 1] = rowptr[i] + len(js)
    return rowptr, np.concatenate(nbr).astype(np.int32), np.concatenate(w).astype(np.float64)
