"""Synthetic structured problems for the SIMP hot path.  Host-side set-up only.

``cantilever2d`` is the problem of /root/reference/sample/optimize/sample_optimize_density_oc.cpp:24-80 at any
resolution (C1 = 60x40 exactly as shipped, C2 = 2000x1000, the 2M-dof headline = 1000x1000);
``heat2d`` is config 3, ``cantilever3d`` configs 4-5 (our x-major hex mesher).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import eqcode, mesher

EQ_PLANESTRAIN, EQ_SOLID, EQ_HEAT = 0, 1, 2
FILTER_DENSITY, FILTER_HEAVISIDE = 0, 1
OPT_OC, OPT_MMA, OPT_CONLIN = 0, 1, 2


@dataclass
class Problem:
    name: str
    eq: int
    coords: np.ndarray          # nnode x dim
    conn: np.ndarray            # nelem x npe (int32)
    fixed: tuple                # (node, dof, value)
    loads: tuple                # (node, dof, value)
    nbrs: tuple                 # filter neighbour lists (rowptr int64, nbr int32, w f64)
    grid: tuple                 # element counts per axis
    filter_kind: int = FILTER_HEAVISIDE
    opt_kind: int = OPT_OC
    # sample_optimize_density_oc.cpp:69-78
    E0: float = 1.0e-4
    E1: float = 2.1e5
    poisson: float = 0.3
    penal: float = 3.0
    weightlimit: float = 0.5
    scale0: float = 1.0e5
    scale1: float = 1.0
    thickness: float = 1.0
    beta0: float = 0.5
    beta_period: int = 40
    cg_itrmax: int = 100000
    cg_eps: float = 1.0e-10
    # OC(n, 0.5, 0, 1e4, 1e-3, 0.15, ...) sample_optimize_density_oc.cpp:80
    oc: tuple = (0.5, 0.0, 1.0e4, 1.0e-3, 0.15)
    # MMA(n,1,a0=1,a={0},c={1e4},d={0},xmin=.01,xmax=1) + SetParameters(1e-5,.1,.2,.5,.7,1.2,1e-6)  ..._mma.cpp:80-85
    mma: tuple = (1.0e-5, 0.1, 0.2, 0.5, 0.7, 1.2, 1.0e-6, 1.0, 0.0, 1.0e4, 0.0, 0.01, 1.0)
    # CONLIN(n,1,a0=1,a={0},c={1e4},d={0},xmin=.01,xmax=1) + SetParameters(0.2, 1e-6)  sample_optimize_density_CONLIN.cpp:80-85
    conlin: tuple = (0.2, 1.0e-6, 1.0, 0.0, 1.0e4, 0.0, 0.01, 1.0)
    s0: float = 0.5
    extra: dict = field(default_factory=dict)

    @property
    def ndof(self):
        return eqcode.ndof(self.eq)

    @property
    def nnode(self):
        return self.coords.shape[0]

    @property
    def nelem(self):
        return self.conn.shape[0]

    def params(self):
        return np.array([self.E0, self.E1, self.poisson, self.penal, self.weightlimit, self.scale0, self.scale1,
                         self.thickness, self.beta0, float(self.beta_period), float(self.cg_itrmax), self.cg_eps])

    def optp(self):
        return np.array({OPT_OC: self.oc, OPT_MMA: self.mma, OPT_CONLIN: self.conlin}[self.opt_kind], dtype=np.float64)

    def free_dofs(self):
        return self.nnode * self.ndof - len(self.fixed[0])


def cantilever2d(nx=60, ny=40, opt_kind=OPT_OC, filter_kind=FILTER_HEAVISIDE, radius=1.5, xr=None) -> Problem:
    """xr = (i0, i1): build only element columns [i0, i1) of the nx x ny problem (one slab, local ids)."""
    lx, ly = float(nx), float(ny)
    coords, conn = mesher.square_mesh(lx, ly, nx, ny, xr)
    fixed = mesher.fixed_list(coords, [0, 1], lambda x: np.abs(x[:, 0]) < 1.0e-5)
    ln, ld, lv = mesher.fixed_list(coords, [1], lambda x: (np.abs(x[:, 0] - lx) < 1.0e-5) & (np.abs(x[:, 1] - ly / 2) < 1.0e-5), -1.0)
    nxl = nx if xr is None else xr[1] - xr[0]
    nbrs = mesher.filter_neighbors_2d(nxl, ny, float(nxl), ly, radius)
    return Problem(f"cantilever2d_{nx}x{ny}", EQ_PLANESTRAIN, coords, conn, fixed, (ln, ld, lv), nbrs, (nxl, ny),
                   filter_kind=filter_kind, opt_kind=opt_kind, extra={"global_grid": (nx, ny), "radius": radius})


def heat2d(nx=64, ny=64, opt_kind=OPT_OC, filter_kind=FILTER_DENSITY, radius=1.5, xr=None) -> Problem:
    """Config 3: Q4 heat conduction, k(rho) = k0 + (k1-k0) rho^p, sink on the middle 10% of the left edge,
    uniform nodal heat load 1/nnode on all free nodes, volume fraction 0.4."""
    lx, ly = float(nx), float(ny)
    coords, conn = mesher.square_mesh(lx, ly, nx, ny, xr)
    fixed = mesher.fixed_list(coords, [0], lambda x: (np.abs(x[:, 0]) < 1.0e-5) & (np.abs(x[:, 1] - ly / 2) <= 0.05 * ly + 1.0e-9))
    nnode = coords.shape[0]
    is_fixed = np.zeros(nnode, bool)
    is_fixed[fixed[0]] = True
    ln = np.nonzero(~is_fixed)[0].astype(np.int32)
    loads = (ln, np.zeros_like(ln), np.full(ln.shape, 1.0 / ((nx + 1) * (ny + 1))))
    nxl = nx if xr is None else xr[1] - xr[0]
    nbrs = mesher.filter_neighbors_2d(nxl, ny, float(nxl), ly, radius)
    return Problem(f"heat2d_{nx}x{ny}", EQ_HEAT, coords, conn, fixed, loads, nbrs, (nxl, ny),
                   filter_kind=filter_kind, opt_kind=opt_kind, E0=1.0e-3, E1=1.0, weightlimit=0.4, scale0=1.0,
                   beta_period=0, extra={"global_grid": (nx, ny), "radius": radius})


def cantilever3d(nx=16, ny=8, nz=8, opt_kind=OPT_OC, filter_kind=FILTER_DENSITY, radius=1.5, xr=None) -> Problem:
    """Configs 4-5: hex8 cantilever, clamp the x=0 face (3 dofs), unit -y load spread over the line x=lx, y=ly/2."""
    lx, ly, lz = float(nx), float(ny), float(nz)
    coords, conn = mesher.box_mesh(lx, ly, lz, nx, ny, nz, xr)
    fixed = mesher.fixed_list(coords, [0, 1, 2], lambda x: np.abs(x[:, 0]) < 1.0e-5)
    sel = lambda x: (np.abs(x[:, 0] - lx) < 1.0e-5) & (np.abs(x[:, 1] - ly / 2) < 1.0e-5)
    ln, ld, lv = mesher.fixed_list(coords, [1], sel, -1.0 / (nz + 1))
    nxl = nx if xr is None else xr[1] - xr[0]
    nbrs = mesher.filter_neighbors_3d(nxl, ny, nz, float(nxl), ly, lz, radius)
    return Problem(f"cantilever3d_{nx}x{ny}x{nz}", EQ_SOLID, coords, conn, fixed, (ln, ld, lv), nbrs, (nxl, ny, nz),
                   filter_kind=filter_kind, opt_kind=opt_kind, beta_period=0, extra={"global_grid": (nx, ny, nz), "radius": radius})


def family_problem(eq, n, opt_kind=OPT_OC, filter_kind=FILTER_DENSITY, radius=1.5) -> Problem:
    """A cantilever (elastic physics) or heat sink (HeatTransfer) on a structured block of ANY element family: `eq` is a
    PF2_EQ_CODE (pansfem2_b200/eqcode.py), n the cell counts.  Clamp / sink on x = 0; unit load on the nodes of the
    x = lx face closest to mid-height (elastic) or a uniform nodal source (heat)."""
    phys, shape, _, _ = eqcode.fields(eq)
    fam = eqcode.SHAPE_NAME[shape]
    coords, conn = mesher.family_mesh(fam, n)
    ndof = eqcode.ndof(eq)
    lx, ly = float(n[0]), float(n[1])
    fixed = mesher.fixed_list(coords, list(range(ndof)), lambda x: np.abs(x[:, 0]) < 1.0e-9)
    if phys == eqcode.PHYS_HEAT:
        is_fixed = np.zeros(coords.shape[0], bool)
        is_fixed[fixed[0]] = True
        ln = np.nonzero(~is_fixed)[0].astype(np.int32)
        loads = (ln, np.zeros_like(ln), np.full(ln.shape, 1.0 / coords.shape[0]))
        kw = dict(E0=1.0e-3, E1=1.0, weightlimit=0.4, scale0=1.0)
    else:
        face = np.abs(coords[:, 0] - lx) < 1.0e-9
        dist = np.where(face, np.abs(coords[:, 1] - ly / 2), np.inf)
        sel = np.nonzero(face & (dist <= dist.min() + 1.0e-9))[0].astype(np.int32)
        loads = (sel, np.ones_like(sel), np.full(sel.shape, -1.0 / len(sel)))
        kw = {}
    nbrs = mesher.filter_neighbors_centroid(mesher.element_centroids(coords, conn), radius)
    return Problem(f"{eqcode.describe(eq)}_{'x'.join(str(v) for v in n)}", eq, coords, conn, fixed, loads, nbrs, tuple(n),
                   filter_kind=filter_kind, opt_kind=opt_kind, beta_period=0, **kw)


@dataclass
class LevelSetProblem:
    """sample/optimize/sample_optimize_levelset.cpp:24-72: Q4 plane-stress cantilever, reaction-diffusion level-set update."""
    coords: np.ndarray
    conn: np.ndarray
    fixed: tuple                # (node, dof, value) of the displacement field
    loads: tuple
    phifixed: np.ndarray        # nodes where phi is held at 0 (the outer boundary)
    grid: tuple
    Vmax: float = 0.5
    tau: float = 2.0e-4
    E0: float = 1.0
    Emin: float = 1.0e-4
    nu: float = 0.3
    nvol: float = 100.0
    dt: float = 0.1
    d: float = -0.02
    p: float = 4.0
    tmax: int = 200

    def prm(self):
        return np.array([self.Vmax, self.tau, self.E0, self.Emin, self.nu, self.nvol, self.dt, self.d, self.p])

    @property
    def nnode(self):
        return self.coords.shape[0]

    @property
    def nelem(self):
        return self.conn.shape[0]


def levelset2d(nx=60, ny=40, **kw) -> LevelSetProblem:
    lx, ly = float(nx), float(ny)
    coords, conn = mesher.square_mesh(lx, ly, nx, ny)
    fixed = mesher.fixed_list(coords, [0, 1], lambda x: np.abs(x[:, 0]) < 1.0e-5)
    loads = mesher.fixed_list(coords, [1], lambda x: (np.abs(x[:, 0] - lx) < 1.0e-5) & (np.abs(x[:, 1] - ly / 2) < 1.0 + 1.0e-5), -1.0)
    edge = (np.abs(coords[:, 0]) < 1.0e-5) | (np.abs(coords[:, 0] - lx) < 1.0e-5) | (np.abs(coords[:, 1]) < 1.0e-5) | (np.abs(coords[:, 1] - ly) < 1.0e-5)
    return LevelSetProblem(coords, conn, fixed, loads, np.nonzero(edge)[0].astype(np.int32), (nx, ny), **kw)
