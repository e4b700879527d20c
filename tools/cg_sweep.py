"""Time the PCG iteration (ms/iteration over a fixed number of iterations) for each SpMV kernel variant / tuning."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pansfem2_b200 import capi, problems  # noqa: E402

args = sys.argv[1:]
kind = args[0] if args else "2d"
dims = [int(a) for a in args[1:]]
if kind == "2d":
    P = problems.cantilever2d(*(dims or [1000, 1000]), opt_kind=problems.OPT_MMA, filter_kind=problems.FILTER_DENSITY)
elif kind == "heat":
    P = problems.heat2d(*(dims or [1024, 1024]))
else:
    P = problems.cantilever3d(*(dims or [96, 48, 48]))
ctx = capi.Context(0)
S = capi.Simp(ctx, P)
A = S.A
rho = ctx.array(np.full(P.nelem, 0.5))
A.assemble(S.mesh, S.dofmap, P.eq, (P.E0, P.E1, P.poisson, P.penal, P.thickness), P.loads, rho=rho)
x = ctx.empty(A.rows)
ITR = 400
bytes_iter = 12 * A.nnz + 112 * A.rows
res = []
configs = [(v, 3, 4) for v in (3, 31, 0)]
for v, st, c in configs:
    try:
        A.set_spmv_variant(v)
    except capi.Pf2Error:
        continue
    A.set_tma_tuning(st, c)
    try:
        A.solve(capi.SOLVER_SCALINGCG, A.device_F(), x, itrmax=20)
    except capi.Pf2Error:
        pass
    ctx.sync()
    ctx.timer_start()
    try:
        A.solve(capi.SOLVER_SCALINGCG, A.device_F(), x, itrmax=ITR)
    except capi.Pf2Error:
        pass
    ms = ctx.timer_stop()
    sp = A.spmv_bench(v, reps=20, flush_l2=False)
    stt = A.solver_stats(reset=True)
    r = dict(variant=v, stages=st, ctas=c, k1=stt["spmv_ms"], k2=stt["update_ms"], k3=stt["pupdate_ms"], ms_per_iter=ms / ITR, pcg_gbs=bytes_iter / (ms / ITR) / 1e6, spmv_b2b_ms=sp,
             spmv_gbs=(12 * A.nnz + 24 * A.rows) / sp / 1e6)
    res.append(r)
    print(r, flush=True)
    if v < 20:
        pass
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"cg_sweep_{P.name}.json"), "w"), indent=1)
