"""torchrun script: partitioned assembly + distributed ScalingCG on N GPUs vs the single-GPU solve of the same problem.
    torchrun --nproc-per-node N tools/dist_solve_check.py [2d NX NY | 3d NX NY NZ] [--big]"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pansfem2_b200 import capi, partition, problems  # noqa: E402

rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
args = [a for a in sys.argv[1:] if not a.startswith("--")]
kind = args[0] if args else "3d"
dims = [int(a) for a in args[1:]]
check = "--big" not in sys.argv
P = problems.cantilever2d(*(dims or [200, 100])) if kind == "2d" else problems.cantilever3d(*(dims or [32, 16, 16]))
rng = np.random.default_rng(3)
rho_g = rng.uniform(0.3, 1.0, P.nelem)
ctx = capi.Context(local_rank)
D = capi.Dist(ctx, rank, world)
S = partition.slab(P, rank, world)
L = S.local
plane_e = int(np.prod(P.grid[1:]))
mesh = capi.Mesh(ctx, L.coords, L.conn)
dm = capi.DofMap(ctx, L.nnode, L.ndof, L.fixed)
A = capi.Csr.pattern(ctx, mesh, dm)
rho = ctx.array(rho_g[S.le0 * plane_e:S.le1 * plane_e])
A.assemble(mesh, dm, L.eq, (L.E0, L.E1, L.poisson, L.penal, L.thickness), L.loads, rho=rho)
D.set_partition(A, S.own_rows, S.row_halo)
if os.environ.get('PF2_P2P', '1') != '0':
    D.enable_p2p(A, S.row_halo)
x = ctx.empty(A.rows)
for rep in range(2):
    ctx.sync(); dist.barrier()
    t0 = time.time()
    ctx.timer_start()
    it, relres = A.solve(capi.SOLVER_SCALINGCG, A.device_F(), x)
    ms = ctx.timer_stop()
    dist.barrier()
xo = x.download()[S.own_rows[0]:S.own_rows[1]]
res = {"rank": rank, "world": world, "problem": P.name, "rows_local": A.rows, "own_rows": S.own_rows, "iters": it, "relres": relres, "ms": ms,
       "ms_per_iter": ms / it}
if check:
    gathered = [None] * world
    dist.all_gather_object(gathered, (S.global_rows, xo))
    if rank == 0:
        mesh_g = capi.Mesh(ctx, P.coords, P.conn)
        dm_g = capi.DofMap(ctx, P.nnode, P.ndof, P.fixed)
        A_g = capi.Csr.pattern(ctx, mesh_g, dm_g)
        A_g.assemble(mesh_g, dm_g, P.eq, (P.E0, P.E1, P.poisson, P.penal, P.thickness), P.loads, rho=ctx.array(rho_g))
        xg = ctx.empty(A_g.rows)
        ctx.timer_start()
        it_g, rr_g = A_g.solve(capi.SOLVER_SCALINGCG, A_g.device_F(), xg)
        ms_g = ctx.timer_stop()
        xg = xg.download()
        xd = np.zeros_like(xg)
        for (lo, hi), part in gathered:
            xd[lo:hi] = part
        res.update(single_gpu_iters=it_g, single_gpu_ms=ms_g, max_rel_diff=float(np.abs(xd - xg).max() / np.abs(xg).max()))
        assert res["max_rel_diff"] < 1e-8, res
        assert abs(it - it_g) <= max(3, it_g // 50), res
if rank == 0:
    print(json.dumps(res, default=float))
dist.destroy_process_group()
