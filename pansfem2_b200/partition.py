"""Row-block (x-slab) partition of the structured problems across the GPUs of one box.

Host-side set-up only.  Numbering is x-major (node id (ny+1)*i + j, resp. ((ny+1)*i + j)*(nz+1) + k), and free dofs
are numbered node-major, so every contiguous range of node planes is a contiguous range of rows:

    rank r owns element planes [e0, e1) and node planes [e0, e1) (+ the last plane nx on the last rank)
    its LOCAL mesh is element planes [e0-1, e1+1) clipped to the domain (one ghost element plane per side) with node
    planes [e0-1, e1+1]; ghost elements are recomputed on both sides, so assembly needs no communication
    rows of owned nodes are complete; x entries of the ghost node planes e0-1 and e1 arrive by halo exchange

A `Slab` carries the local Problem plus the owned ranges and the contiguous halo ranges (dof rows and elements).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import mesher, problems


@dataclass
class Slab:
    rank: int
    nranks: int
    local: problems.Problem       # local mesh / BCs / loads / filter lists
    e0: int                       # owned element planes [e0, e1) (global x index)
    e1: int
    le0: int                      # local element planes [le0, le1)
    le1: int
    own_rows: tuple               # (lo, hi) in local row numbering
    row_halo: tuple               # (sendL_off, recvL_off, cntL, sendR_off, recvR_off, cntR) local rows
    own_elems: tuple              # (lo, hi) local element ids
    elem_halo: tuple              # same 6-tuple for element fields
    own_nodes: tuple              # (lo, hi) local node ids
    global_rows: tuple            # (lo, hi): where the owned rows sit in the global numbering
    n2g_local: np.ndarray         # local nodetoglobal (nnode_local x ndof), -1 = Dirichlet


def _dofmap(nnode, ndof, fixed):
    n2g = np.zeros(nnode * ndof, np.int64)
    n2g[np.asarray(fixed[0], np.int64) * ndof + np.asarray(fixed[1], np.int64)] = -1
    free = n2g != -1
    n2g[free] = np.arange(int(free.sum()))
    return n2g.reshape(nnode, ndof)


def _check_radius(radius, hx=1.0):
    """One ghost element plane per side holds every filter neighbour of an owned element only while the radius stays below two
    element widths; a wider filter would silently truncate the neighbour lists next to the cut."""
    if int(np.floor(radius / hx + 1.0e-12)) > 1:
        raise ValueError(f"filter radius {radius} reaches past the single ghost element plane of a slab (element width {hx}): "
                         "the row-block partition supports radius < 2 element widths")
    return radius


def slab_from_factory(factory, grid, ndof, rank: int, nranks: int) -> Slab:
    """Build one slab WITHOUT materialising the global mesh: `factory(xr=(i0, i1))` returns the Problem restricted to element
    planes [i0, i1) with local ids (problems.cantilever2d / heat2d / cantilever3d all take `xr`).  Used for the big configs;
    `global_rows` is not known locally and is left at (-1, -1)."""
    nx = grid[0]
    assert nranks <= nx
    plane_e = int(np.prod(grid[1:]))
    plane_n = int(np.prod([g + 1 for g in grid[1:]]))
    e0, e1 = rank * nx // nranks, (rank + 1) * nx // nranks
    le0, le1 = max(e0 - 1, 0), min(e1 + 1, nx)
    L = factory(xr=(le0, le1))
    _check_radius(L.extra.get("radius", 1.5))
    on0, on1 = e0, (e1 if rank < nranks - 1 else nx + 1)
    own_node_lo, own_node_hi = (on0 - le0) * plane_n, (on1 - le0) * plane_n
    ln = np.asarray(L.loads[0], np.int64)
    keep = (ln >= own_node_lo) & (ln < own_node_hi)
    L.loads = (np.asarray(L.loads[0])[keep], np.asarray(L.loads[1])[keep], np.asarray(L.loads[2])[keep])
    L.name = f"{L.name}_slab{rank}of{nranks}"
    n2g = _dofmap(L.nnode, ndof, L.fixed)
    return _finish(L, n2g, rank, nranks, e0, e1, le0, le1, on0, on1, plane_e, plane_n, (-1, -1))


def _finish(local, n2g, rank, nranks, e0, e1, le0, le1, on0, on1, plane_e, plane_n, global_rows) -> Slab:
    def rows_of_planes(p0, p1):
        """local row range [lo, hi) of node planes [p0, p1) (global plane indices); an empty range sits where it would start"""
        a, b = (p0 - le0) * plane_n, (p1 - le0) * plane_n
        blk = n2g[a:b].ravel()
        blk = blk[blk >= 0]
        if blk.size == 0:
            k = int((n2g[:a].ravel() >= 0).sum())
            return k, k
        return int(blk.min()), int(blk.max()) + 1

    own_lo, own_hi = rows_of_planes(on0, on1)
    sendL = recvL = sendR = recvR = (0, 0)
    if rank > 0:
        sendL, recvL = rows_of_planes(e0, e0 + 1), rows_of_planes(e0 - 1, e0)
    if rank < nranks - 1:
        sendR, recvR = rows_of_planes(e1 - 1, e1), rows_of_planes(e1, e1 + 1)
    # one count per side serves both directions of the exchange, so the plane sent and the plane received must carry the same number
    # of rows.  That only fails when an exchanged plane holds Dirichlet dofs - a rank whose single element plane touches the clamped
    # face - which would leave the neighbours waiting for different amounts of data: refuse it here.
    for side, (snd, rcv) in (("left", (sendL, recvL)), ("right", (sendR, recvR))):
        if snd[1] - snd[0] != rcv[1] - rcv[0]:
            raise ValueError(f"rank {rank} of {nranks}: the {side} halo would send {snd[1] - snd[0]} rows and receive {rcv[1] - rcv[0]} "
                             "(a boundary-condition plane is being exchanged): use fewer ranks, at least two element planes per rank")
    row_halo = (sendL[0], recvL[0], sendL[1] - sendL[0], sendR[0], recvR[0], sendR[1] - sendR[0])
    el = lambda p: (p - le0) * plane_e
    elem_halo = (el(e0), el(e0 - 1) if rank > 0 else 0, plane_e if rank > 0 else 0,
                 el(e1 - 1), el(e1) if rank < nranks - 1 else 0, plane_e if rank < nranks - 1 else 0)
    return Slab(rank, nranks, local, e0, e1, le0, le1, (own_lo, own_hi), row_halo, (el(e0), el(e1)), elem_halo,
                ((on0 - le0) * plane_n, (on1 - le0) * plane_n), global_rows, n2g.astype(np.int32))


def slab(P: problems.Problem, rank: int, nranks: int) -> Slab:
    nx = P.grid[0]
    assert nranks <= nx, "more ranks than element planes"
    plane_e = int(np.prod(P.grid[1:]))
    plane_n = int(np.prod([g + 1 for g in P.grid[1:]]))
    ndof = P.ndof
    e0, e1 = rank * nx // nranks, (rank + 1) * nx // nranks
    le0, le1 = max(e0 - 1, 0), min(e1 + 1, nx)
    node_lo, node_hi = le0 * plane_n, (le1 + 1) * plane_n
    coords = P.coords[node_lo:node_hi].copy()
    conn = (P.conn[le0 * plane_e:le1 * plane_e].astype(np.int64) - node_lo).astype(np.int32)
    # Dirichlet conditions on local nodes
    fn = np.asarray(P.fixed[0], np.int64)
    keep = (fn >= node_lo) & (fn < node_hi)
    fixed = ((fn[keep] - node_lo).astype(np.int32), np.asarray(P.fixed[1])[keep].astype(np.int32), np.asarray(P.fixed[2])[keep])
    # owned node planes
    on0, on1 = e0, (e1 if rank < nranks - 1 else nx + 1)
    # loads only on owned nodes (each load is applied exactly once)
    ln = np.asarray(P.loads[0], np.int64)
    keep = (ln >= on0 * plane_n) & (ln < on1 * plane_n)
    loads = ((ln[keep] - node_lo).astype(np.int32), np.asarray(P.loads[1])[keep].astype(np.int32), np.asarray(P.loads[2])[keep])
    # filter lists of the local grid (identical to the global lists for every owned element)
    lgrid = (le1 - le0,) + tuple(P.grid[1:])
    radius = _check_radius(P.extra.get("radius", 1.5))
    if len(P.grid) == 2:
        nbrs = mesher.filter_neighbors_2d(lgrid[0], lgrid[1], float(lgrid[0]), float(lgrid[1]), radius)
    else:
        nbrs = mesher.filter_neighbors_3d(lgrid[0], lgrid[1], lgrid[2], float(lgrid[0]), float(lgrid[1]), float(lgrid[2]), radius)
    local = problems.Problem(f"{P.name}_slab{rank}of{nranks}", P.eq, coords, conn, fixed, loads, nbrs, lgrid,
                             filter_kind=P.filter_kind, opt_kind=P.opt_kind, E0=P.E0, E1=P.E1, poisson=P.poisson, penal=P.penal,
                             weightlimit=P.weightlimit, scale0=P.scale0, scale1=P.scale1, thickness=P.thickness, beta0=P.beta0,
                             beta_period=P.beta_period, cg_itrmax=P.cg_itrmax, cg_eps=P.cg_eps, oc=P.oc, mma=P.mma, conlin=P.conlin, s0=P.s0)
    n2g = _dofmap(coords.shape[0], ndof, fixed)
    # global position of the owned rows
    gn2g = _dofmap(P.nnode, ndof, P.fixed)
    gblk = gn2g[on0 * plane_n:on1 * plane_n].ravel()
    gblk = gblk[gblk >= 0]
    global_rows = (int(gblk.min()), int(gblk.max()) + 1) if gblk.size else (0, 0)
    return _finish(local, n2g, rank, nranks, e0, e1, le0, le1, on0, on1, plane_e, plane_n, global_rows)


# ------------------------------------------------------------------------------------------------------------------
# numpy emulation of the distributed Jacobi-PCG (used by the CPU gloo tests to validate the partition / halo logic;
# `comm` provides allreduce(np.ndarray) -> np.ndarray and sendrecv(send_left, send_right) -> (from_left, from_right))
# ------------------------------------------------------------------------------------------------------------------
def pcg_partitioned(matvec, diag, b, S: Slab, comm, itrmax=100000, eps=1e-10):
    lo, hi = S.own_rows
    sL, rL, cL, sR, rR, cR = S.row_halo
    n = b.shape[0]

    def halo(v):
        from_left, from_right = comm.sendrecv(v[sL:sL + cL].copy() if cL else None, v[sR:sR + cR].copy() if cR else None)
        if cL:
            v[rL:rL + cL] = from_left
        if cR:
            v[rR:rR + cR] = from_right

    x = np.zeros(n)
    r = np.zeros(n)
    z = np.zeros(n)
    p = np.zeros(n)
    r[lo:hi] = b[lo:hi]
    z[lo:hi] = r[lo:hi] / diag[lo:hi]
    p[lo:hi] = z[lo:hi]
    bb, rho = comm.allreduce(np.array([r[lo:hi] @ r[lo:hi], z[lo:hi] @ r[lo:hi]]))
    halo(p)
    for k in range(itrmax):
        y = matvec(p)
        pAp = comm.allreduce(np.array([p[lo:hi] @ y[lo:hi]]))[0]
        alpha = rho / pAp
        x[lo:hi] += alpha * p[lo:hi]
        r[lo:hi] -= alpha * y[lo:hi]
        z[lo:hi] = r[lo:hi] / diag[lo:hi]
        zr, rr = comm.allreduce(np.array([z[lo:hi] @ r[lo:hi], r[lo:hi] @ r[lo:hi]]))
        beta = zr / rho
        rho = zr
        p[lo:hi] = beta * p[lo:hi] + z[lo:hi]
        halo(p)
        if np.sqrt(rr) < eps * np.sqrt(bb):
            return x, k + 1
    return x, itrmax


def pcg_partitioned_single_reduction(matvec, diag, b, S: Slab, comm, itrmax=100000, eps=1e-10):
    """Host statement of csrc/dist.cu: solve_dist_cg1 (pf2_csr_set_cg_variant(A, 1)): the Chronopoulos-Gear form of the same Jacobi-PCG with
    ONE allreduce per iteration -- {w.u, u.r, r.r} together -- and the halo exchange of u instead of p.  Returns (x, iterations, allreduces)."""
    lo, hi = S.own_rows
    sL, rL, cL, sR, rR, cR = S.row_halo
    n = b.shape[0]
    own = slice(lo, hi)

    def halo(v):
        from_left, from_right = comm.sendrecv(v[sL:sL + cL].copy() if cL else None, v[sR:sR + cR].copy() if cR else None)
        if cL:
            v[rL:rL + cL] = from_left
        if cR:
            v[rR:rR + cR] = from_right

    x, r, u, p, s = (np.zeros(n) for _ in range(5))
    r[own] = b[own]
    u[own] = r[own] / diag[own]
    halo(u)
    w = matvec(u)
    delta, bb, gamma, rr = comm.allreduce(np.array([w[own] @ u[own], b[own] @ b[own], u[own] @ r[own], r[own] @ r[own]]))
    sums = 1
    alpha, beta = gamma / delta, 0.0
    for k in range(itrmax):
        p[own] = u[own] + beta * p[own]
        s[own] = w[own] + beta * s[own]
        x[own] += alpha * p[own]
        r[own] -= alpha * s[own]
        u[own] = r[own] / diag[own]
        halo(u)
        w = matvec(u)
        delta, gamma_new, rr = comm.allreduce(np.array([w[own] @ u[own], u[own] @ r[own], r[own] @ r[own]]))
        sums += 1
        if np.sqrt(rr) < eps * np.sqrt(bb):
            return x, k + 1, sums
        beta = gamma_new / gamma
        alpha = gamma_new / (delta - beta * gamma_new / alpha)
        gamma = gamma_new
    return x, itrmax, sums

