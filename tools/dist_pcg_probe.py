"""torchrun probe: time per PCG iteration of the row-partitioned ScalingCG on a bench workload's slabs, for every variant of the loop.
    torchrun --nproc-per-node N tools/dist_pcg_probe.py c2 [iterations]
Variants (switched inside one process): three-kernel loop with the halo wait deferred into the product (default) / at the end of the
p-update, with L2-resident loads on / off, and the persistent kernel."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pansfem2_b200 import capi, partition  # noqa: E402

rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
itr = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
dims, ndof, nelem_g, _ = bench.global_sizes(name)
ctx = capi.Context(local_rank)
D = capi.Dist(ctx, rank, world)
S = partition.slab_from_factory(lambda xr: bench.make_problem(name, xr=xr), dims, ndof, rank, world)
L = S.local
mesh = capi.Mesh(ctx, L.coords, L.conn)
dm = capi.DofMap(ctx, L.nnode, L.ndof, L.fixed)
A = capi.Csr.pattern(ctx, mesh, dm)
A.assemble(mesh, dm, L.eq, (L.E0, L.E1, L.poisson, L.penal, L.thickness), L.loads, rho=ctx.array(np.full(L.nelem, 0.5)))
D.set_partition(A, S.own_rows, S.row_halo)
D.enable_p2p(A, S.row_halo)
x = ctx.empty(A.rows)
out = {"workload": name, "world": world, "rows_local": A.rows, "nnz_local": A.nnz, "iterations": itr}
variants = [("3k", 0, {}), ("3k_defer", 0, {"PF2_HALO_DEFER": "1"}), ("3k_l2", 0, {"PF2_SELL_L2_MB": "115"}), ("persistent", 1, {})]
if os.environ.get("PROBE_ONLY"):
    variants = [v for v in variants if v[0] in os.environ["PROBE_ONLY"].split(",")]
out["sell_unroll"] = os.environ.get("PF2_SELL_UNROLL", "6")
for tag, mode, env in variants:
    for k in ("PF2_HALO_DEFER", "PF2_SELL_L2_MB"):
        os.environ.pop(k, None)
    os.environ.update(env)
    A.set_pcg_mode(mode)
    best = None
    for rep in range(3):
        ctx.sync(); dist.barrier()
        ctx.timer_start()
        try:
            A.solve(capi.SOLVER_SCALINGCG, A.device_F(), x, itrmax=itr)
        except capi.Pf2Error as e:
            if e.code != capi.E_NOCONV:
                raise
        ms = ctx.timer_stop()
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = t.item() if best is None else min(best, t.item())
    out[tag] = round(best / itr, 5)
    if mode == 1:
        st = A.pcg_stats()
        out["persistent_detail"] = {k: round(st[k], 5) for k in ("product_ms", "update_ms", "pupdate_ms", "product_wait_ms", "update_wait_ms", "pupdate_wait_ms")} | {"grid": st["grid"], "solves": st["solves"]}
if rank == 0:
    print(json.dumps(out))
D.release_p2p(A)
dist.destroy_process_group()
