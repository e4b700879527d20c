"""Short workload for ncu: build a structured problem, assemble, run `itr` PCG iterations (no convergence expected).
Usage: python tools/ncu_target.py [2d|heat|3d] dims... [--itr N]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pansfem2_b200 import capi, problems  # noqa: E402

args = sys.argv[1:]
itr = 12
if "--itr" in args:
    i = args.index("--itr")
    itr = int(args[i + 1])
    args = args[:i] + args[i + 2:]
mf = "--matrix-free" in args
args = [a for a in args if a != "--matrix-free"]
kind = args[0] if args else "2d"
dims = [int(a) for a in args[1:]]
if kind == "2d":
    P = problems.cantilever2d(*(dims or [1000, 1000]), opt_kind=problems.OPT_MMA, filter_kind=problems.FILTER_DENSITY)
elif kind == "heat":
    P = problems.heat2d(*(dims or [1024, 1024]))
else:
    P = problems.cantilever3d(*(dims or [96, 48, 48]))
ctx = capi.Context(0)
S = capi.Simp(ctx, P, matrix_free=mf)
rho = ctx.array(np.full(P.nelem, 0.5))
S.A.assemble(S.mesh, S.dofmap, P.eq, (P.E0, P.E1, P.poisson, P.penal, P.thickness), P.loads, rho=rho)
x = ctx.empty(S.A.rows)
try:
    S.A.solve(capi.SOLVER_SCALINGCG, S.A.device_F(), x, itrmax=itr)
except capi.Pf2Error as e:
    print("expected:", str(e)[:80])
print("launches", ctx.launch_count())
