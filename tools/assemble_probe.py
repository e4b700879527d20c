"""Numeric assembly on the B200: the row-gather kernel (default for 2-D selections) against the scatter kernel (PF2_ASSEMBLE_SCATTER=1),
device time per assembly, run-to-run reproducibility and agreement of the two.  Usage: python tools/assemble_probe.py [out.json]
(spawns itself once per mode: the switch is read once per process)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [("Q4 plane strain 1000x1000 (specialised)", "ps", (1000, 1000)), ("Q4 heat 1024x1024 (specialised)", "heat", (1024, 1024)),
         ("T3 plane stress 1000x500", "t3", (1000, 500)), ("Q4 SRI 1000x500", "sri", (1000, 500)), ("Q8 plane strain Gauss9 600x300", "q8", (600, 300)),
         ("Hex8 solid 96x48x48 (specialised; gather = PF2_ASSEMBLE_GATHER3D=1: 32-node tiles, one thread per dof row)", "h8", (96, 48, 48))]


if os.environ.get("ASM_ONLY"):
    CASES = [c for c in CASES if c[1] in os.environ["ASM_ONLY"].split(",")]


def worker(out_npz):
    from pansfem2_b200 import capi, mesher
    from pansfem2_b200 import eqcode as ec
    eqs = dict(ps=ec.eq_code(ec.PHYS_PLANESTRAIN), heat=ec.eq_code(ec.PHYS_HEAT), t3=ec.eq_code(ec.PHYS_PLANESTRESS, ec.SHAPE_T3),
               sri=ec.eq_code(ec.PHYS_PLANESTRAIN_SRI, ec.SHAPE_Q4), h8=ec.eq_code(ec.PHYS_SOLID), q8=ec.eq_code(ec.PHYS_PLANESTRAIN, ec.SHAPE_Q8, ec.QUAD_G9SQ))
    ctx = capi.Context(0)
    res, keep = [], {}
    for label, key, n in CASES:
        eq = eqs[key]
        coords, conn = mesher.family_mesh(ec.SHAPE_NAME[ec.fields(eq)[1]], n)
        ndof = ec.ndof(eq)
        fixed = mesher.fixed_list(coords, list(range(ndof)), lambda x: np.abs(x[:, 0]) < 1e-9, value=0.01)
        mesh, dm = capi.Mesh(ctx, coords, conn), capi.DofMap(ctx, coords.shape[0], ndof, fixed)
        A = capi.Csr.pattern(ctx, mesh, dm)
        rho = ctx.array(np.random.default_rng(1).uniform(0.2, 1.0, conn.shape[0]))
        loads = (np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0))
        prm = (1e-4, 2.1e5, 0.3, 3.0, 1.0)
        A.assemble(mesh, dm, eq, prm, loads, rho=rho)
        _, _, d1, F1 = A.download()
        A.assemble(mesh, dm, eq, prm, loads, rho=rho)
        _, _, d2, F2 = A.download()
        reps = 10
        ctx.timer_start()
        for _ in range(reps):
            A.assemble(mesh, dm, eq, prm, loads, rho=rho)
        ms = ctx.timer_stop() / reps
        res.append(dict(label=label, nelem=int(conn.shape[0]), nnz=int(A.nnz), assemble_ms=round(ms, 4), gelem_per_s=round(conn.shape[0] / ms / 1e6, 3),
                        reproducible=bool(np.array_equal(d1, d2) and np.array_equal(F1, F2))))
        keep[key + "_data"], keep[key + "_F"] = d2[::97].copy(), F2[::97].copy()
        for o in (A, dm, mesh):
            o.close()
    np.savez(out_npz, **keep)
    print(json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        worker(sys.argv[2])
        sys.exit(0)
    out = {}
    for mode in ("gather", "scatter"):
        env = dict(os.environ)
        env.pop("PF2_ASSEMBLE_SCATTER", None)
        env["PF2_ASSEMBLE_GATHER3D"] = "1"
        if mode == "scatter":
            env["PF2_ASSEMBLE_SCATTER"] = "1"
        r = subprocess.run([sys.executable, __file__, "--worker", f"/tmp/asm_{mode}.npz"], env=env, capture_output=True, text=True, check=True)
        out[mode] = json.loads(r.stdout.strip().split("\n")[-1])
    g, s = np.load("/tmp/asm_gather.npz"), np.load("/tmp/asm_scatter.npz")
    out["max_rel_diff_gather_vs_scatter"] = {k: float(np.abs(g[k] - s[k]).max() / max(np.abs(s[k]).max(), 1e-300)) for k in g.files}
    for a, b in zip(out["gather"], out["scatter"]):
        a["speedup_vs_scatter"] = round(b["assemble_ms"] / a["assemble_ms"], 2)
    print(json.dumps(out, indent=1))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
