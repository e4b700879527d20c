"""Throughput of the generic element kernels on the B200: symbolic pattern, numeric assembly, the
sensitivity pass and the SpMV the resulting matrix gets, per element family.  Usage: python tools/family_probe.py [out.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pansfem2_b200 import capi, mesher  # noqa: E402
from pansfem2_b200 import eqcode as ec  # noqa: E402

CASES = [
    ("Q4 plane strain (specialised kernel)", ec.eq_code(ec.PHYS_PLANESTRAIN), (1000, 500)),
    ("Q4 plane strain via the generic template", ec.eq_code(ec.PHYS_PLANESTRAIN, ec.SHAPE_Q4, ec.QUAD_G9SQ), (1000, 500)),
    ("Q4 SRI", ec.eq_code(ec.PHYS_PLANESTRAIN_SRI, ec.SHAPE_Q4), (1000, 500)),
    ("T3 plane stress", ec.eq_code(ec.PHYS_PLANESTRESS, ec.SHAPE_T3), (1000, 500)),
    ("T6 plane strain, Gauss3", ec.eq_code(ec.PHYS_PLANESTRAIN, ec.SHAPE_T6, ec.QUAD_G3TRI), (500, 250)),
    ("Q8 plane strain, Gauss9", ec.eq_code(ec.PHYS_PLANESTRAIN, ec.SHAPE_Q8, ec.QUAD_G9SQ), (600, 300)),
    ("T3 heat", ec.eq_code(ec.PHYS_HEAT, ec.SHAPE_T3), (1000, 500)),
    ("Hex8 solid (specialised kernel)", ec.eq_code(ec.PHYS_SOLID), (96, 48, 48)),
    ("Hex8 solid, Gauss27 (generic)", ec.eq_code(ec.PHYS_SOLID, ec.SHAPE_HEX8, ec.QUAD_G27CUBE), (96, 48, 48)),
    ("Tet4 solid", ec.eq_code(ec.PHYS_SOLID, ec.SHAPE_TET4), (64, 32, 32)),
    ("Hex20 solid, Gauss27", ec.eq_code(ec.PHYS_SOLID, ec.SHAPE_HEX20, ec.QUAD_G27CUBE), (48, 24, 24)),
]


def main():
    ctx = capi.Context(0)
    rows = []
    for label, eq, n in CASES:
        shape = ec.fields(eq)[1]
        t0 = time.time()
        coords, conn = mesher.family_mesh(ec.SHAPE_NAME[shape], n)
        ndof = ec.ndof(eq)
        fixed = mesher.fixed_list(coords, list(range(ndof)), lambda x: np.abs(x[:, 0]) < 1e-9)
        t_mesh = time.time() - t0
        mesh = capi.Mesh(ctx, coords, conn)
        dm = capi.DofMap(ctx, coords.shape[0], ndof, fixed)
        ctx.sync(); t0 = time.time()
        A = capi.Csr.pattern(ctx, mesh, dm)
        ctx.sync(); t_pat = time.time() - t0
        rho = ctx.array(np.full(conn.shape[0], 0.5))
        loads = (np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0))
        prm = (1e-4, 2.1e5, 0.3, 3.0, 1.0)
        A.assemble(mesh, dm, eq, prm, loads, rho=rho)
        reps = 5
        ctx.timer_start()
        for _ in range(reps):
            A.assemble(mesh, dm, eq, prm, loads, rho=rho)
        t_asm = ctx.timer_stop() / reps
        u = ctx.array(np.random.default_rng(0).uniform(-1e-3, 1e-3, coords.shape[0] * ndof))
        dfd = ctx.empty(conn.shape[0])
        capi.compliance_sens_device(mesh, eq, u, rho, (1e-4, 2.1e5, 0.3, 3.0, 1.0, 1e5), dfd)
        ctx.timer_start()
        for _ in range(reps):
            capi.compliance_sens_device(mesh, eq, u, rho, (1e-4, 2.1e5, 0.3, 3.0, 1.0, 1e5), dfd)
        t_sens = ctx.timer_stop() / reps
        ms = A.spmv_bench(0, 20, True)
        nnz = A.nnz
        row = dict(selection=ec.describe(eq), label=label, cells=list(n), nelem=int(conn.shape[0]), rows=int(A.rows), nnz=int(nnz),
                   mesh_host_s=round(t_mesh, 2), pattern_ms=round(1e3 * t_pat, 1), assemble_ms=round(t_asm, 3),
                   melem_per_s_assemble=round(conn.shape[0] / t_asm / 1e3, 1), sens_ms=round(t_sens, 3),
                   spmv_ms=round(ms, 4), spmv_algorithmic_GBps=round((12.0 * nnz + 24.0 * A.rows) / ms / 1e6, 1))
        rows.append(row)
        print(json.dumps(row), flush=True)
        for o in (A, dm, mesh):
            o.close()
    if len(sys.argv) > 1:
        json.dump(rows, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
