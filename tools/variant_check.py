"""Check one SpMV kernel variant (plain + fused-dot inside a PCG solve) on one small problem against numpy."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pansfem2_b200 import capi, problems  # noqa: E402

variant, prob = int(sys.argv[1]), sys.argv[2]
P = {"2d": lambda: problems.cantilever2d(40, 30), "heat": lambda: problems.heat2d(33, 17), "3d": lambda: problems.cantilever3d(6, 5, 4)}[prob]()
ctx = capi.Context(0)
S = capi.Simp(ctx, P)
rho = ctx.array(np.random.default_rng(1).uniform(0.3, 1, P.nelem))
S.A.assemble(S.mesh, S.dofmap, P.eq, (P.E0, P.E1, P.poisson, P.penal, P.thickness), P.loads, rho=rho)
indptr, indices, data, F = S.A.download()
try:
    S.A.set_spmv_variant(variant)
except capi.Pf2Error as e:
    print(variant, prob, "n/a")
    sys.exit(0)
x = np.random.default_rng(2).uniform(-1, 1, S.A.rows)
import scipy.sparse as sp
M = sp.csr_matrix((data, indices, indptr), shape=(S.A.rows, S.A.rows))
y = S.A.spmv_host(x)
err = np.abs(y - M @ x).max() / np.abs(M @ x).max()
print(variant, prob, "spmv err", err, flush=True)
t = time.time()
xs, it, rr = S.A.solve_host(capi.SOLVER_SCALINGCG, F, raise_noconv=False)
print(variant, prob, "solve iters", it, "relres", rr, "s", time.time() - t, flush=True)
