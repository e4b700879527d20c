"""Hex20 numeric assembly on the B200: the warp-per-element kernel (default) against the thread-per-(element, node) kernel
(PF2_HEX20_WARP=0): device time per assembly of a 48x24x24 mesh (27 648 elements, Gauss27Cubic) and agreement of the two.
Usage: python tools/hex20_probe.py [out.json]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pansfem2_b200 import capi, mesher  # noqa: E402
from pansfem2_b200 import eqcode as ec  # noqa: E402

n = tuple(int(v) for v in os.environ.get("HEX20_N", "48,24,24").split(","))
eq = ec.eq_code(ec.PHYS_SOLID, ec.SHAPE_HEX20, ec.QUAD_G27CUBE)
ctx = capi.Context(0)
coords, conn = mesher.family_mesh("Hex20", n)
fixed = mesher.fixed_list(coords, [0, 1, 2], lambda x: np.abs(x[:, 0]) < 1e-9, value=0.01)
mesh, dm = capi.Mesh(ctx, coords, conn), capi.DofMap(ctx, coords.shape[0], 3, fixed)
A = capi.Csr.pattern(ctx, mesh, dm)
rho = ctx.array(np.random.default_rng(1).uniform(0.2, 1.0, conn.shape[0]))
loads = (np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0))
prm = (1e-4, 2.1e5, 0.3, 3.0, 1.0)
out = {"mesh": list(n), "nelem": int(conn.shape[0]), "rows": int(A.rows), "nnz": int(A.nnz)}
keep = {}
for mode, sw in (("warp_per_element", "1"), ("thread_per_element_node", "0")):
    os.environ["PF2_HEX20_WARP"] = sw
    A.assemble(mesh, dm, eq, prm, loads, rho=rho)
    _, _, data, F = A.download()
    reps = 5
    ctx.timer_start()
    for _ in range(reps):
        A.assemble(mesh, dm, eq, prm, loads, rho=rho)
    ms = ctx.timer_stop() / reps
    keep[mode] = (data, F)
    out[mode] = {"assemble_ms": round(ms, 4), "melem_per_s": round(conn.shape[0] / ms / 1e3, 2)}
d1, F1 = keep["warp_per_element"]
d0, F0 = keep["thread_per_element_node"]
out["max_rel_diff_data"] = float(np.abs(d1 - d0).max() / np.abs(d0).max())
out["max_rel_diff_F"] = float(np.abs(F1 - F0).max() / max(np.abs(F0).max(), 1e-300))
out["speedup"] = round(out["thread_per_element_node"]["assemble_ms"] / out["warp_per_element"]["assemble_ms"], 2)
print(json.dumps(out))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
