"""Advection-diffusion time stepping at scale on the B200 (rotating cone of sample_advectiondiffusion_dynamic.cpp on an nx x ny Q4 grid):
device time of pf2_advdiff_assemble, and of a whole step (assembly + BiCGSTAB + disassembly).
Usage: python tools/advection_probe.py [nx ny steps]   -> one JSON line, also written to gpurun_out/advection_probe.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pansfem2_b200 import capi, mesher  # noqa: E402
from pansfem2_b200 import eqcode as ec  # noqa: E402

nx, ny, steps = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1000, 1000, 5)
coords, conn = mesher.square_mesh(1.0, 1.0, nx, ny)
edge = np.nonzero((coords[:, 0] < 1e-9) | (coords[:, 0] > 1 - 1e-9) | (coords[:, 1] < 1e-9) | (coords[:, 1] > 1 - 1e-9))[0].astype(np.int32)
r = np.sqrt((coords[:, 0] - 0.5) ** 2 + (coords[:, 1] - 0.75) ** 2)
T0 = np.where(r <= 0.25, 0.5 * (np.cos(4.0 * np.pi * r) + 1.0), 0.0)
cg = coords[conn].mean(axis=1)
vel = np.stack([-(cg[:, 1] - 0.5), cg[:, 0] - 0.5], axis=1)
terms = ec.ADV_ADVECTION | ec.ADV_DIFFUSION | ec.ADV_SUPG | ec.ADV_MASS | ec.ADV_MASS_SUPG
eq = ec.eq_code(ec.PHYS_ADVDIFF, ec.SHAPE_Q4, ec.QUAD_G4SQ, terms)
dt, theta = 0.5 / max(nx, ny), 0.5           # CFL ~ 0.35 at the rim

ctx = capi.Context(0)
mesh = capi.Mesh(ctx, coords, conn)
dmap = capi.DofMap(ctx, len(coords), 1, (edge, np.zeros_like(edge), np.zeros(len(edge))))
K = capi.Csr.pattern(ctx, mesh, dmap)
veld, T, x = ctx.array(vel.ravel()), ctx.array(T0), ctx.empty(K.rows)
prm = (0.0, 0.0, 0.0, 1.0 / dt, theta, 1.0 - theta)
K.advdiff_assemble(mesh, dmap, eq, prm, vel=veld, T=T)      # warm-up
ctx.sync()
reps = 10
ctx.timer_start()
for _ in range(reps):
    K.advdiff_assemble(mesh, dmap, eq, prm, vel=veld, T=T)
asm_ms = ctx.timer_stop() / reps
rows = []
for s in range(steps):
    ctx.timer_start()
    K.advdiff_assemble(mesh, dmap, eq, prm, vel=veld, T=T)
    it, relres = K.solve(capi.SOLVER_BICGSTAB, K.device_F(), x)
    dmap.disassemble(x, T)
    rows.append(dict(step=s, ms=round(ctx.timer_stop(), 3), bicgstab_iters=it, relres=relres))
Tn = T.download()
out = dict(grid=[nx, ny], elements=int(len(conn)), rows=int(K.rows), nnz=int(K.nnz), assemble_ms=round(asm_ms, 4),
           assemble_elements_per_s=round(len(conn) / (asm_ms * 1e-3), 0), steps=rows, T_min=float(Tn.min()), T_max=float(Tn.max()))
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "advection_probe.json"), "w").write(json.dumps(out) + "\n")
