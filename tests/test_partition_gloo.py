"""Multi-process (gloo, world_size 2, 3 and 4: interior ranks with two neighbours) tests of the row-block partition logic on CPU: local meshes with ghost planes,
owned row ranges, contiguous halo ranges, loads applied once.  Each rank assembles ITS slab with the oracle, the ranks run
the partitioned Jacobi-PCG (numpy emulation of csrc/dist.cu: halo exchange of p + allreduces) over torch.distributed/gloo,
and the gathered solution must equal the single-domain oracle solve."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import portlib as orc
from pansfem2_b200 import partition, problems


class GlooComm:
    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def allreduce(self, a):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
        dist.all_reduce(t)
        return t.numpy()

    def sendrecv(self, to_left, to_right):
        reqs, from_left, from_right = [], None, None
        if self.rank > 0 and to_left is not None:
            from_left = torch.zeros(len(to_left), dtype=torch.float64)
            reqs += [dist.isend(torch.from_numpy(to_left), self.rank - 1), dist.irecv(from_left, self.rank - 1)]
        if self.rank < self.world - 1 and to_right is not None:
            from_right = torch.zeros(len(to_right), dtype=torch.float64)
            reqs += [dist.isend(torch.from_numpy(to_right), self.rank + 1), dist.irecv(from_right, self.rank + 1)]
        for r in reqs:
            r.wait()
        return (from_left.numpy() if from_left is not None else None, from_right.numpy() if from_right is not None else None)


def _make(kind):
    if kind == "2d":
        return problems.cantilever2d(24, 10)
    if kind == "heat":
        return problems.heat2d(12, 10)
    return problems.cantilever3d(7, 4, 2)


def _worker(rank, world, port, kind, out_dir, single_reduction=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = _make(kind)
        rng = np.random.default_rng(5)
        rho_global = rng.uniform(0.2, 1.0, P.nelem)
        S = partition.slab(P, rank, world)
        L = S.local
        plane_e = int(np.prod(P.grid[1:]))
        rho_local = rho_global[S.le0 * plane_e:S.le1 * plane_e]
        Emod = L.E1 * rho_local ** L.penal + L.E0 * (1 - rho_local ** L.penal)
        So, n2g, ufix, _ = orc.assemble(L.eq, L.coords, L.conn, L.fixed, L.loads, Emod, L.poisson, L.thickness)
        assert np.array_equal(n2g, S.n2g_local)
        indptr, indices, data, F = So.arrays()
        A = sp.csr_matrix((data, indices, indptr), shape=(So.rows, So.rows))
        if single_reduction:      # csrc/dist.cu: solve_dist_cg1 -- one allreduce per iteration
            x, its, sums = partition.pcg_partitioned_single_reduction(lambda v: A @ v, A.diagonal(), F, S, GlooComm(rank, world))
            assert sums == its + 1
        else:
            x, its = partition.pcg_partitioned(lambda v: A @ v, A.diagonal(), F, S, GlooComm(rank, world))
        lo, hi = S.own_rows
        np.save(os.path.join(out_dir, f"x_{rank}.npy"), x[lo:hi])
        np.save(os.path.join(out_dir, f"meta_{rank}.npy"), np.array([S.global_rows[0], S.global_rows[1], its]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("kind,world,single_reduction", [("2d", 2, False), ("3d", 2, False), ("heat", 3, False), ("2d", 4, False),
                                                         ("2d", 4, True), ("heat", 3, True), ("3d", 3, True)])
def test_partitioned_pcg_matches_single_domain(tmp_path, kind, world, single_reduction):
    """single_reduction: the opt-in recurrences of solve_dist_cg1 -- same solution and iteration count as the reference's ScalingCG"""
    mp.spawn(_worker, args=(world, _free_port(), kind, str(tmp_path), single_reduction), nprocs=world, join=True)
    P = _make(kind)
    rng = np.random.default_rng(5)
    rho = rng.uniform(0.2, 1.0, P.nelem)
    Emod = P.E1 * rho ** P.penal + P.E0 * (1 - rho ** P.penal)
    So, *_ = orc.assemble(P.eq, P.coords, P.conn, P.fixed, P.loads, Emod, P.poisson, P.thickness)
    F = So.arrays()[3]
    xref, it_ref, relres = So.solve(1, F)
    x = np.zeros(So.rows)
    covered = np.zeros(So.rows, bool)
    for r in range(world):
        lo, hi, its = np.load(tmp_path / f"meta_{r}.npy").astype(int)
        x[lo:hi] = np.load(tmp_path / f"x_{r}.npy")
        assert not covered[lo:hi].any()
        covered[lo:hi] = True
        assert abs(its - it_ref) <= max(2, it_ref // 50)
    assert covered.all()                                   # every global row is owned by exactly one rank
    assert np.abs(x - xref).max() < 1e-9 * np.abs(xref).max()


def test_too_many_ranks_is_refused():
    """9 element planes on 8 ranks leave rank 0 with the clamped plane only: its neighbour would send a full plane and receive none."""
    P = problems.cantilever3d(9, 4, 2)
    with pytest.raises(ValueError, match="fewer ranks"):
        for r in range(8):
            partition.slab(P, r, 8)


def test_slab_ranges_cover_elements_and_rows():
    P = problems.cantilever3d(18, 4, 2)
    for world in (1, 2, 4, 8):
        rows, elems = [], []
        for r in range(world):
            S = partition.slab(P, r, world)
            rows.append(S.global_rows)
            elems.append((S.e0, S.e1))
            sL, rL, cL, sR, rR, cR = S.row_halo
            assert (cL == 0) == (r == 0) and (cR == 0) == (r == world - 1)
            if r > 0:
                Sl = partition.slab(P, r - 1, world)
                assert Sl.row_halo[5] == cL                 # what I send left is what my left neighbour receives from its right
        assert rows[0][0] == 0 and rows[-1][1] == P.free_dofs()
        assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
        assert elems[0][0] == 0 and elems[-1][1] == P.grid[0]


def test_slab_refuses_a_filter_radius_wider_than_the_ghost_plane():
    """One ghost element plane per side holds the filter neighbours of an owned element only for radius < 2 element widths."""
    import pytest
    from pansfem2_b200 import partition, problems
    P = problems.cantilever2d(24, 8, radius=2.5)
    with pytest.raises(ValueError, match="ghost element plane"):
        partition.slab(P, 0, 2)
    assert partition.slab(problems.cantilever2d(24, 8, radius=1.5), 1, 2).local.nelem == 13 * 8
