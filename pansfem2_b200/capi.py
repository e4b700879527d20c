"""ctypes binding of libpansfem2_b200.so (C ABI: include/pansfem2_b200.h).

Plumbing for tests and bench.py only - a C++ user binds the same ABI through the header mirror under
pansfem2_b200/src.  There is no CPU fallback: if the shared library is missing or no CUDA device is present every
entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import eqcode
from .eqcode import eq_code  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PF2_LIB") or os.path.join(HERE, "libpansfem2_b200.so")      # PF2_LIB: an experimental build (tuning runs only)

EQ_PLANESTRAIN, EQ_SOLID, EQ_HEAT = 0, 1, 2
SOLVER_CG, SOLVER_SCALINGCG, SOLVER_ILU0CG = 0, 1, 2
SOLVER_BICGSTAB, SOLVER_BICGSTAB2, SOLVER_SCALINGBICGSTAB, SOLVER_ILU0BICGSTAB = 3, 4, 5, 6
FILTER_DENSITY, FILTER_HEAVISIDE, FILTER_SENS_SIGMUND, FILTER_SENS_BORRVALL = 0, 1, 2, 3
OPT_OC, OPT_MMA, OPT_CONLIN = 0, 1, 2
E_NOCONV = 4

_lib = None


class Pf2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pansfem2_b200 error {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing - build it with `python -m pansfem2_b200.build` "
                               "(pansfem2_b200 has no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.pf2_last_error.restype = C.c_char_p
        _lib.pf2_version.restype = C.c_char_p
    return _lib


def _ck(rc, allow=()):
    if rc != 0 and rc not in allow:
        raise Pf2Error(rc, lib().pf2_last_error().decode())
    return rc


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


class DeviceArray:
    """A device buffer owned through pf2_malloc / pf2_free."""

    def __init__(self, ctx, count, dtype=np.float64):
        self.ctx, self.count, self.dtype = ctx, int(count), np.dtype(dtype)
        self.ptr = C.c_void_p()
        _ck(lib().pf2_malloc(ctx.h, C.c_size_t(self.count * self.dtype.itemsize), C.byref(self.ptr)))

    @classmethod
    def from_host(cls, ctx, arr):
        arr = np.ascontiguousarray(arr)
        d = cls(ctx, arr.size, arr.dtype)
        d.upload(arr)
        return d

    def upload(self, arr):
        arr = np.ascontiguousarray(arr, dtype=self.dtype)
        assert arr.size == self.count
        _ck(lib().pf2_memcpy_h2d(self.ctx.h, self.ptr, arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.nbytes)))

    def download(self):
        out = np.empty(self.count, self.dtype)
        _ck(lib().pf2_memcpy_d2h(self.ctx.h, out.ctypes.data_as(C.c_void_p), self.ptr, C.c_size_t(out.nbytes)))
        return out

    def free(self):
        if self.ptr and self.ctx.h:       # (a context that was closed first took its allocations' stream with it)
            lib().pf2_free(self.ctx.h, self.ptr)
        self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    def __init__(self, device=0, stream=None):
        self.h = C.c_void_p()
        _ck(lib().pf2_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self.h)))

    def sync(self):
        _ck(lib().pf2_ctx_sync(self.h))

    def launch_count(self):
        v = C.c_longlong(0)
        _ck(lib().pf2_ctx_launch_count(self.h, C.byref(v)))
        return v.value

    def device_info(self):
        sm, ma, mi, mem = C.c_int(0), C.c_int(0), C.c_int(0), C.c_size_t(0)
        _ck(lib().pf2_ctx_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), total_mem=mem.value)

    def timer_start(self):
        _ck(lib().pf2_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double(0)
        _ck(lib().pf2_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        _ck(lib().pf2_flush_l2(self.h))

    def array(self, arr):
        return DeviceArray.from_host(self, arr)

    def empty(self, count, dtype=np.float64):
        return DeviceArray(self, count, dtype)

    def element_matrix(self, eq, xe, E, V=0.3, t=1.0):
        xe = _f64(xe)
        m = xe.shape[0] * eqcode.ndof(eq)
        Ke = np.zeros((m, m))
        _ck(lib().pf2_element_matrix(self.h, eq, _p(xe, np.float64), C.c_double(E), C.c_double(V), C.c_double(t), _p(Ke, np.float64)))
        return Ke

    def element_matrix_d(self, eq, xe, D, t=1.0):
        """PF2_PHYS_PLANE_D* selections (PlaneStiffness*, Homogenization.h:141-280): D is the 3 x 3 constitutive matrix."""
        xe, D = _f64(xe), _f64(D).reshape(9)
        m = xe.shape[0] * 2
        Ke = np.zeros((m, m))
        _ck(lib().pf2_element_matrix_d(self.h, eq, _p(xe, np.float64), _p(D, np.float64), C.c_double(t), _p(Ke, np.float64)))
        return Ke

    def close(self):
        if self.h:
            lib().pf2_ctx_destroy(self.h)
            self.h = C.c_void_p()


class Mesh:
    def __init__(self, ctx, coords, conn):
        coords, conn = _f64(coords), _i32(conn)
        self.ctx, self.nnode, self.dim = ctx, coords.shape[0], coords.shape[1]
        self.nelem, self.npe = conn.shape
        self.h = C.c_void_p()
        _ck(lib().pf2_mesh_create(ctx.h, self.dim, self.nnode, _p(coords, np.float64), self.npe, self.nelem, _p(conn, np.int32), C.byref(self.h)))

    @classmethod
    def on_nodes(cls, base, conn):
        """A second element list (edges, a sub-region) over the nodes of `base`; shares its coordinates on the device."""
        conn = _i32(conn)
        m = cls.__new__(cls)
        m.ctx, m.nnode, m.dim = base.ctx, base.nnode, base.dim
        m.nelem, m.npe = conn.shape
        m.h = C.c_void_p()
        _ck(lib().pf2_mesh_create_on_nodes(base.h, m.npe, m.nelem, _p(conn, np.int32), C.byref(m.h)))
        return m

    def close(self):
        if self.h:
            lib().pf2_mesh_destroy(self.h)
            self.h = C.c_void_p()


class DofMap:
    def __init__(self, ctx, nnode, ndof, fixed):
        fn, fd, fv = _i32(fixed[0]), _i32(fixed[1]), _f64(fixed[2])
        self.ctx, self.nnode, self.ndof = ctx, nnode, ndof
        self.h = C.c_void_p()
        k = C.c_int(0)
        _ck(lib().pf2_dofmap_create(ctx.h, nnode, ndof, len(fn), _p(fn, np.int32), _p(fd, np.int32), _p(fv, np.float64), C.byref(k), C.byref(self.h)))
        self.kdegree = k.value

    def get(self):
        out = np.zeros((self.nnode, self.ndof), np.int32)
        _ck(lib().pf2_dofmap_get(self.h, _p(out, np.int32)))
        return out

    def disassemble(self, x_dev, u_dev):
        _ck(lib().pf2_disassemble(self.h, x_dev.ptr, u_dev.ptr))

    def close(self):
        if self.h:
            lib().pf2_dofmap_destroy(self.h)
            self.h = C.c_void_p()


class Csr:
    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle
        r, nnz = C.c_int(0), C.c_longlong(0)
        _ck(lib().pf2_csr_info(self.h, C.byref(r), C.byref(nnz)))
        self.rows, self.nnz = r.value, nnz.value

    @classmethod
    def pattern(cls, ctx, mesh, dofmap):
        h = C.c_void_p()
        _ck(lib().pf2_csr_pattern(ctx.h, mesh.h, dofmap.h, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def upload(cls, ctx, indptr, indices, data):
        indptr, indices, data = _i32(indptr), _i32(indices), _f64(data)
        h = C.c_void_p()
        _ck(lib().pf2_csr_upload(ctx.h, len(indptr) - 1, _p(indptr, np.int32), _p(indices, np.int32), _p(data, np.float64), C.byref(h)))
        return cls(ctx, h)

    def download(self):
        indptr, indices = np.zeros(self.rows + 1, np.int64), np.zeros(self.nnz, np.int32)
        data, F = np.zeros(self.nnz), np.zeros(self.rows)
        _ck(lib().pf2_csr_download(self.h, _p(indptr, np.int64), _p(indices, np.int32), _p(data, np.float64), _p(F, np.float64)))
        return indptr, indices, data, F

    def assemble(self, mesh, dofmap, eq, params, loads, modulus=None, rho=None):
        """params = (E0, E1, poisson, p, thickness); modulus / rho are DeviceArrays."""
        ln, ld, lv = _i32(loads[0]), _i32(loads[1]), _f64(loads[2])
        prm = (C.c_double * 5)(*params)
        _ck(lib().pf2_assemble(self.h, mesh.h, dofmap.h, eq, modulus.ptr if modulus is not None else None,
                               rho.ptr if rho is not None else None, prm, len(ln), _p(ln, np.int32), _p(ld, np.int32), _p(lv, np.float64)))

    def advdiff_assemble(self, mesh, dofmap, eq, prm, vel=None, T=None):
        """pf2_advdiff_assemble: prm = (ax, ay, k, cm, ck, cf); vel (nelem*2) / T (nnode) are DeviceArrays or None."""
        p6 = (C.c_double * 6)(*prm)
        _ck(lib().pf2_advdiff_assemble(self.h, mesh.h, dofmap.h, eq, vel.ptr if vel is not None else None, p6, T.ptr if T is not None else None))

    def matrix_free(self, mesh, dofmap, eq):
        """Opt into the matrix-free operator (uniform structured Q4 / hex8 meshes); takes effect from the next assemble."""
        _ck(lib().pf2_csr_matrix_free(self.h, mesh.h, dofmap.h, eq))

    def spmv_host(self, x):
        x = _f64(x)
        y = np.zeros(self.rows)
        _ck(lib().pf2_spmv_host(self.h, _p(x, np.float64), _p(y, np.float64)))
        return y

    def set_spmv_variant(self, variant):
        _ck(lib().pf2_spmv_set_variant(self.h, int(variant)))

    def set_tma_tuning(self, stages, ctas_per_sm):
        _ck(lib().pf2_spmv_set_tma_tuning(self.h, int(stages), int(ctas_per_sm)))

    def spmv_bench(self, variant=0, reps=20, flush_l2=True):
        ms = C.c_double(0)
        _ck(lib().pf2_spmv_bench(self.h, variant, reps, int(flush_l2), C.byref(ms)))
        return ms.value

    def solve_host(self, solver, b, itrmax=100000, eps=1e-10, raise_noconv=True):
        b = _f64(b)
        x = np.zeros(self.rows)
        it, rr = C.c_int(0), C.c_double(0)
        rc = _ck(lib().pf2_solve_host(self.h, solver, _p(b, np.float64), _p(x, np.float64), int(itrmax), C.c_double(eps), C.byref(it), C.byref(rr)),
                 allow=() if raise_noconv else (E_NOCONV,))
        return x, it.value, rr.value

    def solve(self, solver, b_dev, x_dev, itrmax=100000, eps=1e-10, warm=False):
        """warm=True: x_dev holds the initial guess (pf2_solve_x0); otherwise the reference's x0 = 0."""
        it, rr = C.c_int(0), C.c_double(0)
        fn = lib().pf2_solve_x0 if warm else lib().pf2_solve
        _ck(fn(self.h, solver, b_dev.ptr if isinstance(b_dev, DeviceArray) else b_dev, x_dev.ptr, int(itrmax), C.c_double(eps),
               C.byref(it), C.byref(rr)))
        return it.value, rr.value

    def set_cg_variant(self, variant):
        """0: the reference's PCG recurrences, 1: single-reduction (partitioned matrix, peer-memory backend), -1: environment"""
        _ck(lib().pf2_csr_set_cg_variant(self.h, int(variant)))

    def set_pcg_mode(self, mode):
        """1: persistent cooperative PCG kernel, 0: three kernels per iteration, -1: environment default (PF2_PCG)."""
        _ck(lib().pf2_csr_set_pcg_mode(self.h, int(mode)))

    def pcg_stats(self):
        st = (C.c_double * 12)()
        _ck(lib().pf2_csr_pcg_stats(self.h, st))
        return dict(kernel_ms=st[0], iters=int(st[1]), solves=int(st[2]), grid=int(st[3]), product_ms=st[4], update_ms=st[5],
                    pupdate_ms=st[6], sell_entries=int(st[7]), product_wait_ms=st[8], update_wait_ms=st[9], pupdate_wait_ms=st[10],
                    single_reduction_solves=int(st[11]))

    def solver_stats(self, reset=False):
        st = (C.c_double * 8)()
        _ck(lib().pf2_csr_solver_stats(self.h, st))
        if reset:
            _ck(lib().pf2_csr_solver_stats_reset(self.h))
        return dict(spmv_ms=st[0], update_ms=st[1], pupdate_ms=st[2], samples=int(st[3]), iters=int(st[4]), variant=int(st[5]),
                    rows=int(st[6]), nnz=int(st[7]))

    def device_F(self):
        p = C.c_void_p()
        _ck(lib().pf2_csr_device_F(self.h, C.byref(p)))
        return p

    def ilu0(self):
        _ck(lib().pf2_ilu0_factor(self.h))
        data = np.zeros(self.nnz)
        _ck(lib().pf2_ilu0_download(self.h, _p(data, np.float64)))
        return data

    def ilu0_solve_host(self, b):
        b = _f64(b)
        x = np.zeros(self.rows)
        _ck(lib().pf2_ilu0_solve_host(self.h, _p(b, np.float64), _p(x, np.float64)))
        return x

    def close(self):
        if self.h:
            lib().pf2_csr_destroy(self.h)
            self.h = C.c_void_p()


class Filter:
    def __init__(self, ctx, kind, rowptr, nbr, w):
        rowptr, nbr, w = _i64(rowptr), _i32(nbr), _f64(w)
        self.ctx, self.n, self.kind = ctx, len(rowptr) - 1, kind
        self.h = C.c_void_p()
        _ck(lib().pf2_filter_create(ctx.h, kind, self.n, _p(rowptr, np.int64), _p(nbr, np.int32), _p(w, np.float64), C.byref(self.h)))

    def set_beta(self, beta):
        _ck(lib().pf2_filter_set_beta(self.h, C.c_double(beta)))

    def apply_host(self, s):
        s = _f64(s)
        rho = np.zeros(self.n)
        _ck(lib().pf2_filter_apply_host(self.h, _p(s, np.float64), _p(rho, np.float64)))
        return rho

    def sens_host(self, s, dfdrho):
        s, dfdrho = _f64(s), _f64(dfdrho)
        out = np.zeros(self.n)
        _ck(lib().pf2_filter_sens_host(self.h, _p(s, np.float64), _p(dfdrho, np.float64), _p(out, np.float64)))
        return out

    def close(self):
        if self.h:
            lib().pf2_filter_destroy(self.h)
            self.h = C.c_void_p()


class OC:
    def __init__(self, ctx, n, iota, lmin, lmax, leps, move):
        self.ctx, self.n = ctx, n
        self.h = C.c_void_p()
        _ck(lib().pf2_oc_create(ctx.h, n, *[C.c_double(v) for v in (iota, lmin, lmax, leps, move)], C.byref(self.h)))

    def is_convergence(self, f):
        v = C.c_int(0)
        _ck(lib().pf2_oc_is_convergence(self.h, C.c_double(f), C.byref(v)))
        return bool(v.value)

    def update_host(self, filt, weightlimit, scale1, x, f, dfdx, dgdx):
        xd, fd, gd = self.ctx.array(_f64(x)), self.ctx.array(_f64(dfdx)), self.ctx.array(_f64(dgdx))
        steps, lam = C.c_int(0), C.c_double(0)
        _ck(lib().pf2_oc_update(self.h, filt.h, C.c_double(weightlimit), C.c_double(scale1), xd.ptr, C.c_double(f), fd.ptr, gd.ptr,
                                C.byref(steps), C.byref(lam)))
        return xd.download(), steps.value, lam.value

    def close(self):
        if self.h:
            lib().pf2_oc_destroy(self.h)
            self.h = C.c_void_p()


class MMA:
    CREATE = "pf2_mma_create"

    def __init__(self, ctx, n, m, a0, a, c, d, xmin, xmax):
        self.ctx, self.n, self.m = ctx, n, m
        a, c, d = _f64(a), _f64(c), _f64(d)
        xmin = _f64(np.broadcast_to(xmin, (n,)))
        xmax = _f64(np.broadcast_to(xmax, (n,)))
        self.h = C.c_void_p()
        _ck(getattr(lib(), self.CREATE)(ctx.h, n, m, C.c_double(a0), _p(a, np.float64), _p(c, np.float64), _p(d, np.float64),
                                        _p(xmin, np.float64), _p(xmax, np.float64), C.byref(self.h)))

    def set_parameters(self, raa0, albefa, move, asyinit, asydecr, asyincr, epsvalue):
        _ck(lib().pf2_mma_set_parameters(self.h, *[C.c_double(v) for v in (raa0, albefa, move, asyinit, asydecr, asyincr, epsvalue)]))

    def is_convergence(self, f):
        v = C.c_int(0)
        _ck(lib().pf2_mma_is_convergence(self.h, C.c_double(f), C.byref(v)))
        return bool(v.value)

    def update_host(self, x, f, dfdx, g, dgdx):
        xd, fd, gd = self.ctx.array(_f64(x)), self.ctx.array(_f64(dfdx)), self.ctx.array(_f64(dgdx).ravel())
        g = _f64(g)
        steps = C.c_int(0)
        _ck(lib().pf2_mma_update(self.h, xd.ptr, C.c_double(f), fd.ptr, _p(g, np.float64), gd.ptr, C.byref(steps)))
        return xd.download(), steps.value

    def close(self):
        if self.h:
            lib().pf2_mma_destroy(self.h)
            self.h = C.c_void_p()


class CONLIN(MMA):
    """CONLIN<T> (CONLIN.h): the MMA handle in CONLIN mode; update_host / is_convergence are inherited."""
    CREATE = "pf2_conlin_create"

    def set_parameters(self, move, epsvalue):
        _ck(lib().pf2_conlin_set_parameters(self.h, C.c_double(move), C.c_double(epsvalue)))


SHAPE_LINE2, SHAPE_LINE3, QUAD_G1LINE, QUAD_G2LINE = 8, 9, 9, 10


def integration_points(mesh, shape, quad, ngauss):
    """x_g of every element of `mesh` (host array nelem x ngauss x 2): where the reference evaluates a load functor."""
    xg = mesh.ctx.empty(mesh.nelem * ngauss * 2)
    _ck(lib().pf2_integration_points(mesh.h, int(shape), int(quad), xg.ptr))
    return xg.download().reshape(mesh.nelem, ngauss, 2)


def load_vector(mesh, dofmap, shape, quad, F_dev, t=1.0, f_const=None, f_gauss=None):
    """pf2_load_vector: adds the surface / body load vector of every element of `mesh` into the device vector F_dev."""
    fc = _f64(f_const) if f_const is not None else None
    fg = mesh.ctx.array(_f64(f_gauss).ravel()) if f_gauss is not None else None
    _ck(lib().pf2_load_vector(mesh.h, dofmap.h, int(shape), int(quad), _p(fc, np.float64) if fc is not None else None,
                              fg.ptr if fg is not None else None, C.c_double(t), F_dev.ptr))


def compliance_sens(mesh, eq, u_dev, rho_dev, params6, want_r=False):
    """params6 = (E0, E1, poisson, p, thickness, scale0).  Returns (f, dfdrho, r or None) as host arrays."""
    ctx = mesh.ctx
    dfd = ctx.empty(mesh.nelem)
    r = ctx.empty(mesh.nnode * eqcode.ndof(eq)) if want_r else None
    f = C.c_double(0)
    prm = (C.c_double * 6)(*params6)
    _ck(lib().pf2_compliance_sens(mesh.h, eq, u_dev.ptr, rho_dev.ptr, prm, C.byref(f), dfd.ptr, r.ptr if r is not None else None))
    return f.value, dfd.download(), (r.download().reshape(mesh.nnode, -1) if r is not None else None)


def compliance_sens_device(mesh, eq, u_dev, rho_dev, params6, dfdrho_dev, r_dev=None):
    """Same pass with caller-owned device outputs and no host read-back (the compliance stays in the context's scalar slot)."""
    prm = (C.c_double * 6)(*params6)
    _ck(lib().pf2_compliance_sens(mesh.h, eq, u_dev.ptr, rho_dev.ptr, prm, None, dfdrho_dev.ptr, r_dev.ptr if r_dev is not None else None))


class Simp:
    """The device-resident design loop for a pansfem2_b200.problems.Problem."""

    def __init__(self, ctx, problem, solver=SOLVER_SCALINGCG, matrix_free=False):
        P = problem
        self.ctx, self.P = ctx, P
        self.mesh = Mesh(ctx, P.coords, P.conn)
        self.dofmap = DofMap(ctx, P.nnode, P.ndof, P.fixed)
        self.A = Csr.pattern(ctx, self.mesh, self.dofmap)
        if matrix_free:
            self.A.matrix_free(self.mesh, self.dofmap, P.eq)
        self.filter = Filter(ctx, P.filter_kind, *P.nbrs)
        ln, ld, lv = _i32(P.loads[0]), _i32(P.loads[1]), _f64(P.loads[2])
        optp, params = _f64(P.optp()), _f64(P.params())
        self.h = C.c_void_p()
        _ck(lib().pf2_simp_create(ctx.h, self.mesh.h, self.dofmap.h, self.A.h, self.filter.h, P.eq, P.opt_kind, _p(optp, np.float64),
                                  _p(params, np.float64), len(ln), _p(ln, np.int32), _p(ld, np.int32), _p(lv, np.float64), C.byref(self.h)))
        _ck(lib().pf2_simp_set_solver(self.h, solver))
        self.set_design(np.full(P.nelem, P.s0))

    def set_design(self, s):
        s = _f64(s)
        _ck(lib().pf2_simp_set_design(self.h, _p(s, np.float64)))

    def reset(self, s=None):
        """Back to design iteration 0 with design s (default: the uniform initial design)."""
        s = _f64(np.full(self.P.nelem, self.P.s0) if s is None else s)
        _ck(lib().pf2_simp_reset(self.h, _p(s, np.float64)))

    def set_warm_start(self, on=True):
        _ck(lib().pf2_simp_set_warm_start(self.h, int(bool(on))))

    @staticmethod
    def _stats(st):
        return dict(f=st[0], g=st[1], converged=bool(st[2]), cg_iters=int(st[3]), cg_relres=st[4], opt_steps=int(st[5]), beta=st[6], k=int(st[7]))

    def iterate(self, check_convergence=True):
        st = (C.c_double * 8)()
        _ck(lib().pf2_simp_iterate(self.h, int(check_convergence), st))
        return self._stats(st)

    def iterate_host(self, s_in, s_out, rho_out, check_convergence=True):
        """End-to-end variant: s_in / s_out / rho_out are host numpy arrays (pinned or not)."""
        st = (C.c_double * 8)()
        _ck(lib().pf2_simp_iterate_host(self.h, int(check_convergence), _p(s_in, np.float64) if s_in is not None else None,
                                        _p(s_out, np.float64), _p(rho_out, np.float64), st))
        return self._stats(st)

    def get(self, want_r=False):
        P = self.P
        s, rho = np.zeros(P.nelem), np.zeros(P.nelem)
        u = np.zeros((P.nnode, P.ndof))
        r = np.zeros((P.nnode, P.ndof)) if want_r else None
        _ck(lib().pf2_simp_get(self.h, _p(s, np.float64), _p(rho, np.float64), _p(u, np.float64), _p(r, np.float64) if want_r else None))
        return dict(s=s, rho=rho, u=u, r=r)

    def phase_ms(self):
        ms = (C.c_double * 6)()
        _ck(lib().pf2_simp_phase_ms(self.h, ms))
        return dict(zip(("filter", "assemble", "solve", "sens", "filter_sens", "update"), list(ms)))

    def close(self):
        if self.h:
            lib().pf2_simp_destroy(self.h)
            self.h = C.c_void_p()
        for o in (self.filter, self.A, self.dofmap, self.mesh):
            o.close()


class LevelSet:
    """The device-resident level-set loop for a pansfem2_b200.problems.LevelSetProblem (sample_optimize_levelset.cpp)."""

    def __init__(self, ctx, P, matrix_free=False):
        self.ctx, self.P = ctx, P
        self.mesh = Mesh(ctx, P.coords, P.conn)
        self.dofmap = DofMap(ctx, P.nnode, 2, P.fixed)
        self.K = Csr.pattern(ctx, self.mesh, self.dofmap)
        if matrix_free:     # the displacement solve applies K matrix-free (plane-stress Q4 on the uniform SquareMesh)
            self.K.matrix_free(self.mesh, self.dofmap, eq_code(eqcode.PHYS_PLANESTRESS, eqcode.SHAPE_Q4, eqcode.QUAD_G4SQ))
        pn = _i32(P.phifixed)
        ln, ld, lv = _i32(P.loads[0]), _i32(P.loads[1]), _f64(P.loads[2])
        prm = _f64(P.prm())
        self.h = C.c_void_p()
        _ck(lib().pf2_levelset_create(ctx.h, self.mesh.h, self.dofmap.h, self.K.h, len(pn), _p(pn, np.int32), _p(prm, np.float64), int(P.tmax),
                                      len(ln), _p(ln, np.int32), _p(ld, np.int32), _p(lv, np.float64), C.byref(self.h)))

    def set_state(self, phi, st):
        phi, st = _f64(phi), _f64(st)
        _ck(lib().pf2_levelset_set_state(self.h, _p(phi, np.float64), _p(st, np.float64)))

    def iterate(self, check_convergence=True):
        st = (C.c_double * 8)()
        _ck(lib().pf2_levelset_iterate(self.h, int(check_convergence), st))
        return dict(objective=st[0], vol=st[1], lam=st[2], converged=bool(st[3]), cg_iters=int(st[4]), cg_relres=st[5], cg_iters_phi=int(st[6]), t=int(st[7]))

    def get(self):
        phi, st, u = np.zeros(self.P.nnode), np.zeros(self.P.nelem), np.zeros((self.P.nnode, 2))
        _ck(lib().pf2_levelset_get(self.h, _p(phi, np.float64), _p(st, np.float64), _p(u, np.float64)))
        return dict(phi=phi, str=st, u=u)

    def close(self):
        if self.h:
            lib().pf2_levelset_destroy(self.h)
            self.h = None
        for o in (self.K, self.dofmap, self.mesh):
            o.close()


class Dist:
    """Row-block partition over the GPUs of one box.  `group` is an initialised torch.distributed process group (any
    backend): it is only used to hand rank 0's NCCL id to the other ranks."""

    def __init__(self, ctx, rank, world):
        import torch
        import torch.distributed as tdist
        self.ctx, self.rank, self.world = ctx, rank, world
        idbuf = (C.c_char * 128)()
        if rank == 0:
            _ck(lib().pf2_dist_unique_id(idbuf))
        obj = [bytes(idbuf.raw) if rank == 0 else None]
        tdist.broadcast_object_list(obj, src=0)
        idbuf.raw = obj[0]
        self.h = C.c_void_p()
        _ck(lib().pf2_dist_create(ctx.h, rank, world, idbuf, C.byref(self.h)))

    def allreduce(self, dev, count):
        _ck(lib().pf2_dist_allreduce_sum(self.h, dev.ptr, int(count)))

    def halo(self, dev, halo6):
        h = (C.c_int * 6)(*[int(v) for v in halo6])
        _ck(lib().pf2_dist_halo(self.h, dev.ptr, h))

    def set_partition(self, A, own_rows, row_halo):
        h = (C.c_int * 6)(*[int(v) for v in row_halo])
        _ck(lib().pf2_csr_set_partition(A.h, self.h, int(own_rows[0]), int(own_rows[1]), h))

    def enable_p2p(self, A, row_halo):
        """Switch the partitioned PCG of matrix A to the peer-memory backend (call after set_partition on every rank)."""
        import torch.distributed as tdist
        hb = (C.c_char * 128)()
        _ck(lib().pf2_csr_p2p_export(A.h, hb))
        cap = C.c_int(0)
        _ck(lib().pf2_csr_pcg_capable(A.h, C.byref(cap)))
        meta = [int(v) for v in row_halo] + [int(A.rows), int(cap.value)]
        gathered = [None] * self.world
        tdist.all_gather_object(gathered, (bytes(hb.raw), meta))
        allh = (C.c_char * (128 * self.world))()
        allh.raw = b"".join(g[0] for g in gathered)
        allm = (C.c_int * (8 * self.world))(*[v for g in gathered for v in g[1]])
        _ck(lib().pf2_csr_p2p_import(A.h, allh, allm))
        tdist.barrier()

    def release_p2p(self, A):
        """Collective: unmap the neighbours' slabs of matrix A on every rank, then synchronise, so that A can be destroyed."""
        import torch.distributed as tdist
        _ck(lib().pf2_csr_p2p_release(A.h))
        tdist.barrier()

    def set_simp_partition(self, simp, slab, n_global_elems):
        """simp: capi.Simp built on slab.local; also partitions its matrix."""
        self.set_partition(simp.A, slab.own_rows, slab.row_halo)
        if os.environ.get("PF2_P2P", "1") != "0" and self.world <= 8:
            self.enable_p2p(simp.A, slab.row_halo)
        h = (C.c_int * 6)(*[int(v) for v in slab.elem_halo])
        _ck(lib().pf2_simp_set_partition(simp.h, self.h, int(slab.own_elems[0]), int(slab.own_elems[1]), h, C.c_longlong(int(n_global_elems))))

    def close(self):
        if self.h:
            lib().pf2_dist_destroy(self.h)
            self.h = C.c_void_p()


def pinned_empty(count, dtype=np.float64):
    """numpy view over cudaHostAlloc'ed memory (for the end-to-end path)."""
    dtype = np.dtype(dtype)
    p = C.c_void_p()
    _ck(lib().pf2_host_alloc(C.c_size_t(count * dtype.itemsize), C.byref(p)))
    buf = (C.c_char * (count * dtype.itemsize)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=count)
    return arr
