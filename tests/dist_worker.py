"""torchrun worker of tests/test_gpu_dist.py: the row-partitioned SIMP design loop on N GPUs vs the single-GPU loop of the same problem.
    torchrun --nproc-per-node N tests/dist_worker.py [2d NX NY | 3d NX NY NZ | heat NX NY] [--iters K] [--mma] [--matrix-free] [--warm] [--big]
Environment: PF2_P2P=0 selects the NCCL backend, PF2_PCG=0 the three-kernel PCG loop, PF2_CG_SINGLE_REDUCTION=1 the single-reduction recurrences.  Rank 0 prints one JSON line and asserts
K (through the objective), u, f and the design after K iterations against the single-GPU loop (1e-8 relative / 1e-6 max-abs)."""
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
argv = sys.argv[1:]
iters = 4
if "--iters" in argv:
    i = argv.index("--iters"); iters = int(argv[i + 1]); argv = argv[:i] + argv[i + 2:]
args = [a for a in argv if not a.startswith("--")]
kind = args[0] if args else "3d"
dims = [int(a) for a in args[1:]]
check = "--big" not in argv
opt = problems.OPT_MMA if "--mma" in argv else problems.OPT_OC
if kind == "2d":
    P = problems.cantilever2d(*(dims or [120, 60]), filter_kind=problems.FILTER_HEAVISIDE, opt_kind=opt)
elif kind == "heat":
    P = problems.heat2d(*(dims or [64, 64]), opt_kind=opt)
else:
    P = problems.cantilever3d(*(dims or [24, 12, 8]), opt_kind=opt)
ctx = capi.Context(local_rank)
D = capi.Dist(ctx, rank, world)
S = partition.slab(P, rank, world)
sim = capi.Simp(ctx, S.local, matrix_free=("--matrix-free" in argv))
D.set_simp_partition(sim, S, P.nelem)
sim.set_warm_start("--warm" in argv)
if "--ilu" in argv:       # block-Jacobi ILU(0) per rank: another preconditioner, the same converged solution
    capi._ck(capi.lib().pf2_simp_set_solver(sim.h, capi.SOLVER_ILU0CG))
hist = []
launches0 = ctx.launch_count()
ctx.sync(); dist.barrier()
t0 = time.time()
for k in range(iters):
    st = sim.iterate(check_convergence=False)
    st["phase_ms"] = sim.phase_ms()
    hist.append(st)
ctx.sync(); dist.barrier()
wall = time.time() - t0
out = sim.get()
res = {"world": world, "problem": P.name, "iters": iters, "wall_s": wall, "f": [h["f"] for h in hist], "cg_iters": [h["cg_iters"] for h in hist],
       "opt_steps": [h["opt_steps"] for h in hist], "phase_ms_last": hist[-1]["phase_ms"], "it_per_s": iters / wall}
res["launches_per_cg_iter"] = (ctx.launch_count() - launches0) / max(1, sum(res["cg_iters"]))
if check:
    gathered = [None] * world
    lo, hi = S.own_elems
    plane_e = int(np.prod(P.grid[1:]))
    nlo, nhi = S.own_nodes
    plane_n = int(np.prod([g + 1 for g in P.grid[1:]]))
    on0 = S.e0
    dist.all_gather_object(gathered, (S.e0 * plane_e, S.e1 * plane_e, out["s"][lo:hi], out["rho"][lo:hi], on0 * plane_n, out["u"][nlo:nhi]))
    if rank == 0:
        ref = capi.Simp(ctx, P, solver=capi.SOLVER_ILU0CG if "--ilu" in argv else capi.SOLVER_SCALINGCG)
        fr = [ref.iterate(check_convergence=False) for _ in range(iters)]
        o = ref.get()
        s_d, rho_d, u_d = np.zeros(P.nelem), np.zeros(P.nelem), np.zeros((P.nnode, P.ndof))
        for a, b, sv, rv, n0, uv in gathered:
            s_d[a:b] = sv; rho_d[a:b] = rv; u_d[n0:n0 + uv.shape[0]] = uv
        res.update(f_single=[h["f"] for h in fr], cg_single=[h["cg_iters"] for h in fr],
                   max_s_diff=float(np.abs(s_d - o["s"]).max()), max_rho_diff=float(np.abs(rho_d - o["rho"]).max()),
                   max_u_rel=float(np.abs(u_d - o["u"]).max() / np.abs(o["u"]).max()),
                   max_f_rel=float(max(abs(a["f"] - b["f"]) / abs(b["f"]) for a, b in zip(hist, fr))),
                   pcg=sim.A.pcg_stats())
        assert res["max_f_rel"] < 1e-8 and res["max_s_diff"] < 1e-6 and res["max_rho_diff"] < 1e-6 and res["max_u_rel"] < 1e-7, res
        assert all(abs(a - b) <= max(3, b // 50) for a, b in zip(res["cg_iters"], res["cg_single"])) or "--warm" in argv or "--ilu" in argv, res
        if "--warm" in argv:          # the single-GPU loop started every solve from 0: the warm partitioned loop needs fewer iterations from k = 1 on
            assert sum(res["cg_iters"][1:]) < sum(res["cg_single"][1:]), res
if rank == 0:
    print(json.dumps(res, default=float))
dist.destroy_process_group()
