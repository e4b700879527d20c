"""Level-set loop at scale on the B200: seconds per design iteration and where they go.  Usage: python tools/levelset_probe.py [nx ny iters]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pansfem2_b200 import capi, problems  # noqa: E402

nx, ny, iters = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1000, 500, 6)
P = problems.levelset2d(nx, ny, tmax=iters + 2)
ctx = capi.Context(0)
t0 = time.time()
L = capi.LevelSet(ctx, P, matrix_free=(os.environ.get("PF2_OPERATOR") == "matrix-free"))
ctx.sync()
setup = time.time() - t0
rows = []
for t in range(iters):
    ctx.sync(); t0 = time.time()
    st = L.iterate()
    ctx.sync()
    rows.append(dict(t=t, seconds=round(time.time() - t0, 4), cg_iters_u=st["cg_iters"], cg_iters_phi=st["cg_iters_phi"], objective=st["objective"], vol=st["vol"]))
    print(json.dumps(rows[-1]), flush=True)
print(json.dumps(dict(grid=[nx, ny], dofs_u=L.K.rows, setup_s=round(setup, 2), mean_s_per_iteration=round(float(np.mean([r["seconds"] for r in rows[1:]])), 4))))
