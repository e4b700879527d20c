"""TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/libpf2oracle.so (plain-C restatement, oracle/pf2_oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpf2oracle.so")

EQ_PLANESTRAIN, EQ_SOLID, EQ_HEAT = 0, 1, 2
FILTER_DENSITY, FILTER_HEAVISIDE = 0, 1
OPT_OC, OPT_MMA, OPT_CONLIN = 0, 1, 2


def ndof_of(eq):
    """dofs per node of an eq code (include/pansfem2_b200.h PF2_EQ_CODE): the physics is the low byte."""
    phys = eq & 0xff
    return 3 if phys == 1 else (1 if phys in (2, 5) else 2)

_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "port"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "pf2_oracle.c")
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
            build()
        _lib = C.CDLL(LIB_PATH)
        for name in ("orc_pattern", "orc_system_from_csr", "orc_ilu0", "orc_mma_create"):
            getattr(_lib, name).restype = C.c_void_p
        _lib.orc_system_nnz.restype = C.c_longlong
        _lib.orc_compliance_sens.restype = C.c_double
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


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def num_threads():
    return lib().orc_num_threads()


def element_matrix(eq, xe, E, V=0.3, t=1.0):
    xe = _f64(xe)
    m = xe.shape[0] * ndof_of(eq)
    Ke = np.zeros((m, m))
    lib().orc_element_matrix(eq, _p(xe, np.float64), C.c_double(E), C.c_double(V), C.c_double(t), _p(Ke, np.float64))
    return Ke


def dofmap(nnode, ndof, fixed):
    fn, fd, fv = _i32(fixed[0]), _i32(fixed[1]), _f64(fixed[2])
    n2g = np.zeros((nnode, ndof), np.int32)
    ufix = np.zeros((nnode, ndof))
    k = lib().orc_dofmap(nnode, ndof, len(fn), _p(fn, np.int32), _p(fd, np.int32), _p(fv, np.float64), _p(n2g, np.int32), _p(ufix, np.float64))
    return k, n2g, ufix


class System:
    def __init__(self, handle):
        self.h = C.c_void_p(handle)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.orc_system_free(self.h)
            self.h = None

    @property
    def rows(self):
        return lib().orc_system_rows(self.h)

    @property
    def nnz(self):
        return lib().orc_system_nnz(self.h)

    def arrays(self):
        n, nnz = self.rows, self.nnz
        indptr, indices, data, F = np.zeros(n + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz), np.zeros(n)
        lib().orc_system_get(self.h, _p(indptr, np.int32), _p(indices, np.int32), _p(data, np.float64), _p(F, np.float64))
        return indptr, indices, data, F

    def spmv(self, x):
        x = _f64(x)
        y = np.zeros(self.rows)
        lib().orc_spmv(self.h, _p(x, np.float64), _p(y, np.float64))
        return y

    def solve(self, kind, b, itrmax=100000, eps=1e-10, M=None):
        """kind 0 CG, 1 ScalingCG, 2 ILU0CG, 3 BiCGSTAB, 4 BiCGSTAB2, 5 ScalingBiCGSTAB, 6 ILU0BiCGSTAB.  Returns (x, iterations, relres)."""
        b = _f64(b)
        x = np.zeros(self.rows)
        relres = C.c_double(0)
        if kind >= 3:
            if kind == 6 and M is None:
                M = self.ilu0()
            it = lib().orc_solve_bicgstab(self.h, M.h if M is not None else None, kind, _p(b, np.float64), int(itrmax), C.c_double(eps),
                                          _p(x, np.float64), C.byref(relres))
            return x, it, relres.value
        if kind == 2 and M is None:
            M = self.ilu0()
        it = lib().orc_solve(self.h, M.h if M is not None else None, kind, _p(b, np.float64), int(itrmax), C.c_double(eps),
                             _p(x, np.float64), C.byref(relres))
        return x, it, relres.value

    def ilu0(self):
        return System(lib().orc_ilu0(self.h))

    def preilu0(self, b):
        b = _f64(b)
        x = np.zeros(self.rows)
        lib().orc_preilu0(self.h, _p(b, np.float64), _p(x, np.float64))
        return x


def system_from_csr(indptr, indices, data):
    indptr, indices, data = _i32(indptr), _i32(indices), _f64(data)
    return System(lib().orc_system_from_csr(len(indptr) - 1, _p(indptr, np.int32), _p(indices, np.int32), _p(data, np.float64)))


def assemble(eq, coords, conn, fixed, loads, Emod, V=0.3, t=1.0):
    """Returns (System, nodetoglobal, ufixed, times)."""
    coords, conn = _f64(coords), _i32(conn)
    nnode, ndof = coords.shape[0], ndof_of(eq)
    k, n2g, ufix = dofmap(nnode, ndof, fixed)
    S = System(lib().orc_pattern(nnode, ndof, conn.shape[1], conn.shape[0], _p(conn, np.int32), _p(n2g, np.int32), k))
    ln, ld, lv = _i32(loads[0]), _i32(loads[1]), _f64(loads[2])
    Emod = _f64(Emod)
    times = (C.c_double * 2)()
    lib().orc_assemble_numeric(S.h, eq, _p(coords, np.float64), conn.shape[0], _p(conn, np.int32), _p(n2g, np.int32),
                               _p(ufix, np.float64), _p(Emod, np.float64), C.c_double(V), C.c_double(t),
                               len(ln), _p(ln, np.int32), _p(ld, np.int32), _p(lv, np.float64), times)
    return S, n2g, ufix, {"element": times[0], "scatter": times[1]}


def filter_apply(kind, nbrs, beta, s):
    rowptr, nbr, w = _i64(nbrs[0]), _i32(nbrs[1]), _f64(nbrs[2])
    s = _f64(s)
    n = len(rowptr) - 1
    rho = np.zeros(n)
    lib().orc_filter_apply(kind, n, _p(rowptr, np.int64), _p(nbr, np.int32), _p(w, np.float64), C.c_double(beta), _p(s, np.float64), _p(rho, np.float64))
    return rho


def filter_sens(kind, nbrs, beta, s, dfdrho):
    rowptr, nbr, w = _i64(nbrs[0]), _i32(nbrs[1]), _f64(nbrs[2])
    s, dfdrho = _f64(s), _f64(dfdrho)
    n = len(rowptr) - 1
    out = np.zeros(n)
    lib().orc_filter_sens(kind, n, _p(rowptr, np.int64), _p(nbr, np.int32), _p(w, np.float64), C.c_double(beta),
                          _p(s, np.float64), _p(dfdrho, np.float64), _p(out, np.float64))
    return out


def oc_update(oc, fkind, nbrs, beta, weightlimit, scale1, x, dfdx, dgdx):
    """oc = (iota, lmin, lmax, leps, move).  Returns (x_new, steps, last_lambda)."""
    rowptr, nbr, w = _i64(nbrs[0]), _i32(nbrs[1]), _f64(nbrs[2])
    x = _f64(x).copy()
    dfdx, dgdx = _f64(dfdx), _f64(dgdx)
    lam = C.c_double(0)
    steps = lib().orc_oc_update(len(x), *[C.c_double(v) for v in oc], fkind, _p(rowptr, np.int64), _p(nbr, np.int32), _p(w, np.float64),
                                C.c_double(beta), C.c_double(weightlimit), C.c_double(scale1), _p(x, np.float64),
                                _p(dfdx, np.float64), _p(dgdx, np.float64), C.byref(lam))
    return x, steps, lam.value


def is_convergence(f, fprev, eps):
    return bool(lib().orc_is_convergence(C.c_double(f), C.c_double(fprev), C.c_double(eps)))


class MMA:
    def __init__(self, n, m, a0, a, c, d, xmin, xmax):
        self.n, self.m = n, m
        a, c, d = _f64(a), _f64(c), _f64(d)
        xmin = _f64(np.broadcast_to(xmin, (n,)))
        xmax = _f64(np.broadcast_to(xmax, (n,)))
        self.h = C.c_void_p(lib().orc_mma_create(n, m, C.c_double(a0), _p(a, np.float64), _p(c, np.float64), _p(d, np.float64),
                                                 _p(xmin, np.float64), _p(xmax, np.float64)))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.orc_mma_free(self.h)
            self.h = None

    def set_parameters(self, raa0, albefa, move, asyinit, asydecr, asyincr, epsvalue=None):
        lib().orc_mma_setparameters(self.h, *[C.c_double(v) for v in (raa0, albefa, move, asyinit, asydecr, asyincr)])

    def set_conlin(self, move):
        lib().orc_mma_set_conlin(self.h, C.c_double(move))

    def update(self, x, dfdx, g, dgdx):
        x = _f64(x).copy()
        dfdx, g, dgdx = _f64(dfdx), _f64(g), _f64(dgdx)
        lib().orc_mma_update(self.h, _p(x, np.float64), _p(dfdx, np.float64), _p(g, np.float64), _p(dgdx, np.float64))
        return x

    def stats(self):
        a, b = C.c_int(0), C.c_int(0)
        lib().orc_mma_stats(self.h, C.byref(a), C.byref(b))
        return a.value, b.value


def sensitivity_filter(kind, nbrs, s, dfds):
    rowptr, nbr, w = _i64(nbrs[0]), _i32(nbrs[1]), _f64(nbrs[2])
    s, dfds = _f64(s), _f64(dfds)
    n = len(rowptr) - 1
    out = np.zeros(n)
    lib().orc_sensitivity_filter(kind, n, _p(rowptr, np.int64), _p(nbr, np.int32), _p(w, np.float64), _p(s, np.float64), _p(dfds, np.float64), _p(out, np.float64))
    return out


def compliance_sens(eq, coords, conn, u, rho, E0, E1, V, t, p, scale0):
    coords, conn, u, rho = _f64(coords), _i32(conn), _f64(u), _f64(rho)
    r = np.zeros_like(u)
    dfdrho = np.zeros(conn.shape[0])
    f = lib().orc_compliance_sens(eq, coords.shape[0], _p(coords, np.float64), conn.shape[0], _p(conn, np.int32), _p(u, np.float64),
                                  _p(rho, np.float64), C.c_double(E0), C.c_double(E1), C.c_double(V), C.c_double(t),
                                  C.c_double(p), C.c_double(scale0), _p(r, np.float64), _p(dfdrho, np.float64))
    return f, r, dfdrho


def simp_run(eq, coords, conn, fixed, loads, filter_kind, nbrs, opt_kind, optp, params, niter, s0, check_convergence=True):
    coords, conn = _f64(coords), _i32(conn)
    nnode, nelem, ndof = coords.shape[0], conn.shape[0], ndof_of(eq)
    fn, fd, fv = _i32(fixed[0]), _i32(fixed[1]), _f64(fixed[2])
    ln, ld, lv = _i32(loads[0]), _i32(loads[1]), _f64(loads[2])
    rowptr, nbr, w = _i64(nbrs[0]), _i32(nbrs[1]), _f64(nbrs[2])
    optp, params = _f64(optp), _f64(params)
    s = _f64(s0).copy()
    rho, u, r = np.zeros(nelem), np.zeros((nnode, ndof)), np.zeros((nnode, ndof))
    hist, phase = np.zeros((niter, 5)), np.zeros(8)
    it = lib().orc_simp_run(eq, nnode, _p(coords, np.float64), nelem, _p(conn, np.int32),
                            len(fn), _p(fn, np.int32), _p(fd, np.int32), _p(fv, np.float64),
                            len(ln), _p(ln, np.int32), _p(ld, np.int32), _p(lv, np.float64),
                            filter_kind, _p(rowptr, np.int64), _p(nbr, np.int32), _p(w, np.float64),
                            opt_kind, _p(optp, np.float64), _p(params, np.float64), niter, int(check_convergence),
                            _p(s, np.float64), _p(rho, np.float64), _p(u, np.float64), _p(r, np.float64),
                            _p(hist, np.float64), _p(phase, np.float64))
    return dict(s=s, rho=rho, u=u, r=r, hist=hist[:it], phase=phase, iters=it)


def levelset_run(coords, conn, fixed, loads, phifixed_nodes, prm, tmax, phi0, str0):
    """The level-set loop of sample/optimize/sample_optimize_levelset.cpp (orc_levelset_run); same contract as reflib.levelset_run."""
    coords, conn = _f64(coords), _i32(conn)
    fn, fd, fv = _i32(fixed[0]), _i32(fixed[1]), _f64(fixed[2])
    ln, ld, lv = _i32(loads[0]), _i32(loads[1]), _f64(loads[2])
    pn = _i32(phifixed_nodes)
    prm = _f64(prm)
    nnode, nelem = coords.shape[0], conn.shape[0]
    phi, st = _f64(phi0).copy(), _f64(str0).copy()
    u, hist = np.zeros((nnode, 2)), np.zeros((tmax, 3))
    conv = C.c_int(0)
    it = lib().orc_levelset_run(nnode, _p(coords, np.float64), nelem, _p(conn, np.int32),
                                len(fn), _p(fn, np.int32), _p(fd, np.int32), _p(fv, np.float64),
                                len(ln), _p(ln, np.int32), _p(ld, np.int32), _p(lv, np.float64),
                                len(pn), _p(pn, np.int32), _p(prm, np.float64), int(tmax),
                                _p(phi, np.float64), _p(st, np.float64), _p(u, np.float64), _p(hist, np.float64), C.byref(conv))
    return dict(hist=hist[:it], phi=phi, str=st, u=u, iters=it, converged=bool(conv.value))


ADV_TERMS = dict(advection=1, diffusion=2, supg=4, shock=8, mass=16, mass_supg=32)
_NPE = {1: 3, 2: 6, 3: 4, 4: 8}


def advdiff_element(shape, quad, terms, xe, ax, ay, k):
    """Sum of the selected Advection.h routines (Advection.h:19-229) on one element <SF, IC>."""
    xe = _f64(xe)
    npe = _NPE[shape]
    Ke = np.zeros((npe, npe))
    lib().orc_advdiff_element(shape, quad, terms, _p(xe, np.float64), C.c_double(ax), C.c_double(ay), C.c_double(k), _p(Ke, np.float64))
    return Ke


def advdiff_system(shape, quad, terms, coords, conn, fixed_nodes, fixed_vals, vel, k, dt=0.0, theta=0.5, Tn=None):
    """dt = 0: K, F of sample_advectiondiffusion_static.cpp; dt > 0: one step of sample_advectiondiffusion_dynamic.cpp.
    Returns (System, nodetoglobal, Tn with the Dirichlet values written)."""
    coords, conn, vel = _f64(coords), _i32(conn), _f64(vel)
    nnode = coords.shape[0]
    fn = _i32(fixed_nodes)
    kdeg, n2g, ufix = dofmap(nnode, 1, (fn, np.zeros_like(fn), _f64(fixed_vals)))
    T = np.zeros(nnode) if Tn is None else _f64(Tn).copy()
    T[fn] = _f64(fixed_vals)                       # SetDirichlet (BoundaryCondition.h:20-25)
    S = System(lib().orc_pattern(nnode, 1, conn.shape[1], conn.shape[0], _p(conn, np.int32), _p(n2g, np.int32), kdeg))
    lib().orc_advdiff_assemble(S.h, shape, quad, terms, _p(coords, np.float64), conn.shape[0], _p(conn, np.int32), _p(n2g, np.int32),
                               _p(vel, np.float64), C.c_double(k), C.c_double(dt), C.c_double(theta), _p(T, np.float64))
    return S, n2g, T


def element_matrix_d(eq, xe, D, t=1.0):
    """PlaneStiffness / PlaneStiffnessBbar / PlaneStiffnessWilsonTaylor (Homogenization.h:141-280): eq with phys 10 / 11 / 12, D 3 x 3."""
    xe, D = _f64(xe), _f64(D).reshape(9)
    m = 2 * xe.shape[0]
    Ke = np.zeros((m, m))
    lib().orc_element_matrix_d(int(eq), _p(xe, np.float64), _p(D, np.float64), C.c_double(t), _p(Ke, np.float64))
    return Ke


def load_ngauss(quad):
    return lib().orc_load_ngauss(int(quad))


def integration_points(shape, quad, xe):
    """x_g = X^T N(r_g) of one element (where the reference evaluates the force functor)."""
    xe = _f64(xe)
    xg = np.zeros((load_ngauss(quad), 2))
    lib().orc_integration_points(shape, quad, _p(xe, np.float64), _p(xg, np.float64))
    return xg


def load_vector(shape, quad, ndof, xe, fg, t=1.0):
    """Fe[npe*ndof] of one element from the force density fg[ngauss][ndof] at its integration points."""
    xe, fg = _f64(xe), _f64(fg)
    Fe = np.zeros(xe.shape[0] * ndof)
    lib().orc_load_vector(shape, quad, ndof, _p(xe, np.float64), _p(fg, np.float64), C.c_double(t), _p(Fe, np.float64))
    return Fe
