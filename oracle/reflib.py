"""TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_ref/libpf2ref.so (the UNMODIFIED reference headers behind
an extern "C" shim, oracle/ref_shim.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product path (pansfem2_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpf2ref.so")

EQ_PLANESTRAIN, EQ_SOLID, EQ_HEAT = 0, 1, 2
FILTER_DENSITY, FILTER_HEAVISIDE = 0, 1
OPT_OC, OPT_MMA, OPT_CONLIN = 0, 1, 2


def ndof_of(eq):
    """dofs per node of an eq code (include/pansfem2_b200.h PF2_EQ_CODE): the physics is the low byte."""
    phys = eq & 0xff
    return 3 if phys == 1 else (1 if phys in (2, 5) else 2)

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_assemble.restype = C.c_void_p
        _lib.ref_system_from_csr.restype = C.c_void_p
        _lib.ref_ilu0.restype = C.c_void_p
        _lib.ref_filter_create.restype = C.c_void_p
        _lib.ref_oc_create.restype = C.c_void_p
        _lib.ref_mma_create.restype = C.c_void_p
        _lib.ref_conlin_create.restype = C.c_void_p
        _lib.ref_advection_system.restype = C.c_void_p
        _lib.ref_advdiff_system.restype = C.c_void_p
        _lib.ref_system_nnz.restype = C.c_longlong
    return _lib


def _p(a, dtype):
    if a is None:
        return None
    assert a.dtype == dtype and a.flags["C_CONTIGUOUS"], (a.dtype, dtype)
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def num_threads() -> int:
    return lib().ref_num_threads()


def set_num_threads(n: int):
    lib().ref_set_num_threads(int(n))


def element_matrix(eq, xe, E, V=0.3, t=1.0):
    xe = _f64(xe)
    npe, dim = xe.shape
    m = npe * ndof_of(eq)
    Ke = np.zeros((m, m))
    lib().ref_element_matrix(eq, dim, npe, _p(xe, np.float64), C.c_double(E), C.c_double(V), C.c_double(t), _p(Ke, np.float64))
    return Ke


class System:
    """Reference CSR<double> + F held on the C++ side."""

    def __init__(self, handle):
        self.h = C.c_void_p(handle)
        self.times = None

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_system_free(self.h)
            self.h = None

    @property
    def rows(self):
        return lib().ref_system_rows(self.h)

    @property
    def nnz(self):
        return lib().ref_system_nnz(self.h)

    def arrays(self):
        n, nnz = self.rows, self.nnz
        indptr = np.zeros(n + 1, np.int32)
        indices = np.zeros(nnz, np.int32)
        data = np.zeros(nnz)
        F = np.zeros(n)
        lib().ref_system_get(self.h, _p(indptr, np.int32), _p(indices, np.int32), _p(data, np.float64), _p(F, np.float64))
        return indptr, indices, data, F

    def nodetoglobal(self, nnode, ndof):
        out = np.zeros((nnode, ndof), np.int32)
        lib().ref_system_nodetoglobal(self.h, _p(out, np.int32))
        return out

    def spmv(self, x, repeat=1):
        x = _f64(x)
        y = np.zeros(self.rows)
        sec = C.c_double(0)
        lib().ref_spmv(self.h, _p(x, np.float64), _p(y, np.float64), repeat, C.byref(sec))
        return y, sec.value

    def solve(self, kind, b, itrmax=100000, eps=1e-10):
        """kind: 0 CG, 1 ScalingCG, 2 ILU0CG, 3 BiCGSTAB, 4 BiCGSTAB2, 5 ScalingBiCGSTAB, 6 ILU0BiCGSTAB.  Returns (x, seconds_solve, seconds_factor)."""
        b = _f64(b)
        x = np.zeros(self.rows)
        sec = (C.c_double * 2)()
        lib().ref_solve(self.h, kind, _p(b, np.float64), itrmax, C.c_double(eps), _p(x, np.float64), sec)
        return x, sec[0], sec[1]

    def ilu0(self):
        return System(lib().ref_ilu0(self.h))

    def preilu0(self, b):
        b = _f64(b)
        x = np.zeros(self.rows)
        lib().ref_preilu0(self.h, _p(b, np.float64), _p(x, np.float64))
        return x


def system_from_csr(indptr, indices, data):
    indptr, indices, data = _i32(indptr), _i32(indices), _f64(data)
    return System(lib().ref_system_from_csr(len(indptr) - 1, _p(indptr, np.int32), _p(indices, np.int32), _p(data, np.float64)))


def assemble(eq, coords, conn, fixed, loads, Emod, V=0.3, t=1.0):
    coords, conn = _f64(coords), _i32(conn)
    fn, fd, fv = _i32(fixed[0]), _i32(fixed[1]), _f64(fixed[2])
    ln, ld, lv = _i32(loads[0]), _i32(loads[1]), _f64(loads[2])
    Emod = _f64(Emod)
    times = (C.c_double * 3)()
    h = lib().ref_assemble(eq, coords.shape[1], coords.shape[0], _p(coords, np.float64), conn.shape[1], conn.shape[0],
                           _p(conn, np.int32), len(fn), _p(fn, np.int32), _p(fd, np.int32), _p(fv, np.float64),
                           len(ln), _p(ln, np.int32), _p(ld, np.int32), _p(lv, np.float64),
                           _p(Emod, np.float64), C.c_double(V), C.c_double(t), times)
    s = System(h)
    s.times = {"element": times[0], "assembling": times[1], "tocsr": times[2]}
    return s


class Filter:
    def __init__(self, kind, rowptr, nbr, w):
        self.rowptr, self.nbr, self.w = _i64(rowptr), _i32(nbr), _f64(w)
        self.n = len(self.rowptr) - 1
        self.kind = kind
        self.h = C.c_void_p(lib().ref_filter_create(kind, self.n, _p(self.rowptr, np.int64), _p(self.nbr, np.int32), _p(self.w, np.float64)))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_filter_free(self.h)
            self.h = None

    def apply(self, beta, s):
        s = _f64(s)
        rho = np.zeros(self.n)
        lib().ref_filter_apply(self.h, C.c_double(beta), _p(s, np.float64), _p(rho, np.float64))
        return rho

    def sens(self, beta, s, dfdrho):
        s, dfdrho = _f64(s), _f64(dfdrho)
        out = np.zeros(self.n)
        lib().ref_filter_sens(self.h, C.c_double(beta), _p(s, np.float64), _p(dfdrho, np.float64), _p(out, np.float64))
        return out


class OC:
    def __init__(self, n, iota, lmin, lmax, leps, move):
        self.n = n
        self.h = C.c_void_p(lib().ref_oc_create(n, C.c_double(iota), C.c_double(lmin), C.c_double(lmax), C.c_double(leps), C.c_double(move)))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_oc_free(self.h)
            self.h = None

    def is_convergence(self, f):
        return bool(lib().ref_oc_isconvergence(self.h, C.c_double(f)))

    def update(self, filt: Filter, beta, weightlimit, scale1, s, f, dfds, g, dgds):
        s = _f64(s).copy()
        dfds, dgds = _f64(dfds), _f64(dgds)
        lib().ref_oc_update(self.h, filt.h, C.c_double(beta), C.c_double(weightlimit), C.c_double(scale1), self.n,
                            _p(s, np.float64), C.c_double(f), _p(dfds, np.float64), C.c_double(g), _p(dgds, np.float64))
        return s


class MMA:
    def __init__(self, n, m, a0, a, c, d, xmin, xmax):
        self.n, self.m = n, m
        a, c, d = _f64(a), _f64(c), _f64(d)
        xmin = _f64(np.broadcast_to(xmin, (n,)))
        xmax = _f64(np.broadcast_to(xmax, (n,)))
        self.h = C.c_void_p(lib().ref_mma_create(n, m, C.c_double(a0), _p(a, np.float64), _p(c, np.float64), _p(d, np.float64),
                                                 _p(xmin, np.float64), _p(xmax, np.float64)))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_mma_free(self.h)
            self.h = None

    def set_parameters(self, raa0, albefa, move, asyinit, asydecr, asyincr, epsvalue):
        lib().ref_mma_setparameters(self.h, *[C.c_double(v) for v in (raa0, albefa, move, asyinit, asydecr, asyincr, epsvalue)])

    def is_convergence(self, f):
        return bool(lib().ref_mma_isconvergence(self.h, C.c_double(f)))

    def update(self, x, f, dfdx, g, dgdx):
        x = _f64(x).copy()
        dfdx, g, dgdx = _f64(dfdx), _f64(g), _f64(dgdx)
        lib().ref_mma_update(self.h, self.n, self.m, _p(x, np.float64), C.c_double(f), _p(dfdx, np.float64), _p(g, np.float64), _p(dgdx, np.float64))
        return x


class CONLIN:
    def __init__(self, n, m, a0, a, c, d, xmin, xmax):
        self.n, self.m = n, m
        a, c, d = _f64(a), _f64(c), _f64(d)
        xmin = _f64(np.broadcast_to(xmin, (n,)))
        xmax = _f64(np.broadcast_to(xmax, (n,)))
        self.h = C.c_void_p(lib().ref_conlin_create(n, m, C.c_double(a0), _p(a, np.float64), _p(c, np.float64), _p(d, np.float64),
                                                    _p(xmin, np.float64), _p(xmax, np.float64)))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_conlin_free(self.h)
            self.h = None

    def set_parameters(self, move, epsvalue):
        lib().ref_conlin_setparameters(self.h, C.c_double(move), C.c_double(epsvalue))

    def is_convergence(self, f):
        return bool(lib().ref_conlin_isconvergence(self.h, C.c_double(f)))

    def update(self, x, f, dfdx, g, dgdx):
        x = _f64(x).copy()
        dfdx, g, dgdx = _f64(dfdx), _f64(g), _f64(dgdx)
        lib().ref_conlin_update(self.h, self.n, self.m, _p(x, np.float64), C.c_double(f), _p(dfdx, np.float64), _p(g, np.float64), _p(dgdx, np.float64))
        return x


def sensitivity_filter(kind, nbrs, s, dfds):
    rowptr, nbr, w = _i64(nbrs[0]), _i32(nbrs[1]), _f64(nbrs[2])
    s, dfds = _f64(s), _f64(dfds)
    n = len(rowptr) - 1
    out = np.zeros(n)
    lib().ref_sensitivity_filter(kind, n, _p(rowptr, np.int64), _p(nbr, np.int32), _p(w, np.float64), _p(s, np.float64), _p(dfds, np.float64), _p(out, np.float64))
    return out


def simp_run(eq, coords, conn, fixed, loads, filter_kind, nbrs, opt_kind, optp, params, niter, s0, check_convergence=True):
    """Drive ref_simp_run (the sample's design loop).  Returns dict(s, rho, u, r, hist, phase, iters)."""
    coords, conn = _f64(coords), _i32(conn)
    nnode, dim = coords.shape
    nelem, npe = conn.shape
    ndof = ndof_of(eq)
    fn, fd, fv = _i32(fixed[0]), _i32(fixed[1]), _f64(fixed[2])
    ln, ld, lv = _i32(loads[0]), _i32(loads[1]), _f64(loads[2])
    rowptr, nbr, w = _i64(nbrs[0]), _i32(nbrs[1]), _f64(nbrs[2])
    optp, params = _f64(optp), _f64(params)
    s = _f64(s0).copy()
    rho = np.zeros(nelem)
    u = np.zeros((nnode, ndof))
    r = np.zeros((nnode, ndof))
    hist = np.zeros((niter, 4))
    phase = np.zeros(8)
    it = lib().ref_simp_run(eq, dim, nnode, _p(coords, np.float64), npe, nelem, _p(conn, np.int32),
                            len(fn), _p(fn, np.int32), _p(fd, np.int32), _p(fv, np.float64),
                            len(ln), _p(ln, np.int32), _p(ld, np.int32), _p(lv, np.float64),
                            filter_kind, _p(rowptr, np.int64), _p(nbr, np.int32), _p(w, np.float64),
                            opt_kind, _p(optp, np.float64), _p(params, np.float64), niter, int(check_convergence),
                            _p(s, np.float64), _p(rho, np.float64), _p(u, np.float64), _p(r, np.float64),
                            _p(hist, np.float64), _p(phase, np.float64))
    return dict(s=s, rho=rho, u=u, r=r, hist=hist[:it], phase=phase, iters=it)


def squaremesh(lx, ly, nx, ny):
    coords = np.zeros(((nx + 1) * (ny + 1), 2))
    conn = np.zeros((nx * ny, 4), np.int32)
    lib().ref_squaremesh(C.c_double(lx), C.c_double(ly), nx, ny, _p(coords, np.float64), _p(conn, np.int32))
    return coords, conn


def levelset_run(coords, conn, fixed, loads, phifixed_nodes, prm, tmax, phi0, str0):
    """sample/optimize/sample_optimize_levelset.cpp through the reference's own routines (oracle/ref_shim.cpp ref_levelset_run).
    prm = (Vmax, tau, E0, Emin, nu, nvol, dt, d, p).  Returns dict(hist[t] = (objective, vol, lambda), phi, str, u, iters, converged)."""
    coords, conn = _f64(coords), _i32(conn)
    fn, fd, fv = _i32(fixed[0]), _i32(fixed[1]), _f64(fixed[2])
    ln, ld, lv = _i32(loads[0]), _i32(loads[1]), _f64(loads[2])
    pn = _i32(phifixed_nodes)
    prm = _f64(prm)
    nnode, nelem = coords.shape[0], conn.shape[0]
    phi, st = _f64(phi0).copy(), _f64(str0).copy()
    u, hist = np.zeros((nnode, 2)), np.zeros((tmax, 3))
    conv = C.c_int(0)
    it = lib().ref_levelset_run(nnode, _p(coords, np.float64), nelem, _p(conn, np.int32),
                                len(fn), _p(fn, np.int32), _p(fd, np.int32), _p(fv, np.float64),
                                len(ln), _p(ln, np.int32), _p(ld, np.int32), _p(lv, np.float64),
                                len(pn), _p(pn, np.int32), _p(prm, np.float64), int(tmax),
                                _p(phi, np.float64), _p(st, np.float64), _p(u, np.float64), _p(hist, np.float64), C.byref(conv))
    return dict(hist=hist[:it], phi=phi, str=st, u=u, iters=it, converged=bool(conv.value))


def advection_system(coords, conn, fixed_nodes, fixed_vals, a=1.0, theta_deg=60.0, k=1.0e-6):
    """K, F of sample/advection/sample_advectiondiffusion_static.cpp (T3; Advection + Diffusion + AdvectionSUPG)."""
    coords, conn = _f64(coords), _i32(conn)
    fn, fv = _i32(fixed_nodes), _f64(fixed_vals)
    return System(lib().ref_advection_system(coords.shape[0], _p(coords, np.float64), conn.shape[0], _p(conn, np.int32), len(fn),
                                             _p(fn, np.int32), _p(fv, np.float64), C.c_double(a), C.c_double(theta_deg), C.c_double(k)))


ADV_TERMS = dict(advection=1, diffusion=2, supg=4, shock=8, mass=16, mass_supg=32)
_NPE = {1: 3, 2: 6, 3: 4, 4: 8}


def advdiff_element(shape, quad, terms, xe, ax, ay, k):
    """Sum of the selected Advection.h routines (Advection.h:19-229) on one element <SF, IC>."""
    xe = _f64(xe)
    npe = _NPE[shape]
    Ke = np.zeros((npe, npe))
    lib().ref_advdiff_element(shape, quad, terms, npe, _p(xe, np.float64), C.c_double(ax), C.c_double(ay), C.c_double(k), _p(Ke, np.float64))
    return Ke


def advdiff_system(shape, quad, terms, coords, conn, fixed_nodes, fixed_vals, vel, k, dt=0.0, theta=0.5, Tn=None):
    """dt = 0: K, F of sample_advectiondiffusion_static.cpp; dt > 0: one step of sample_advectiondiffusion_dynamic.cpp."""
    coords, conn, vel = _f64(coords), _i32(conn), _f64(vel)
    fn, fv = _i32(fixed_nodes), _f64(fixed_vals)
    Tn = None if Tn is None else _f64(Tn)
    return System(lib().ref_advdiff_system(shape, quad, terms, coords.shape[0], _p(coords, np.float64), conn.shape[1], conn.shape[0],
                                           _p(conn, np.int32), len(fn), _p(fn, np.int32), _p(fv, np.float64), _p(vel, np.float64),
                                           C.c_double(k), C.c_double(dt), C.c_double(theta), _p(Tn, np.float64)))


def plane_d_element(shape, quad, quad2, mode, xe, D, t=1.0):
    """PlaneStiffness (mode 0) / PlaneStiffnessBbar (1; quad = ICD, quad2 = ICV) / PlaneStiffnessWilsonTaylor (2) of Homogenization.h."""
    xe, D = _f64(xe), _f64(D).reshape(9)
    m = 2 * xe.shape[0]
    Ke = np.zeros((m, m))
    lib().ref_plane_d_element(shape, quad, quad2, mode, xe.shape[0], _p(xe, np.float64), _p(D, np.float64), C.c_double(t), _p(Ke, np.float64))
    return Ke


def load_vector(kind, shape, quad, xe, coef, t=1.0):
    """The reference's PlaneStrain / PlaneStress SurfaceForce / BodyForce (kind 0 / 1) or HeatTransferSurfaceFlux (kind 2) on one element with the
    affine force density f_i(x) = coef[3i] + coef[3i+1] x + coef[3i+2] y.  shape 8 / 9: 2Line / 3Line, else the area shapes of eqcode."""
    xe, coef = _f64(xe), _f64(coef)
    ndof = 1 if kind == 2 else 2
    Fe = np.zeros(xe.shape[0] * ndof)
    lib().ref_load_vector(kind, shape, quad, xe.shape[0], _p(xe, np.float64), _p(coef, np.float64), C.c_double(t), _p(Fe, np.float64))
    return Fe
