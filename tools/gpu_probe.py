"""GPU probe: assembly rate, SpMV kernel-variant sweep, one ScalingCG solve and a few design iterations on a structured
problem.  Writes gpurun_out/probe_<name>.json.  Usage: python tools/gpu_probe.py [2d NX NY | heat NX NY | 3d NX NY NZ] [--iters K]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pansfem2_b200 import capi, problems  # noqa: E402


def main():
    args = sys.argv[1:]
    kind = args[0] if args else "2d"
    iters = 2
    if "--iters" in args:
        i = args.index("--iters")
        iters = int(args[i + 1])
        args = args[:i] + args[i + 2:]
    dims = [int(a) for a in args[1:] if a.isdigit()]
    t0 = time.time()
    if kind == "2d":
        P = problems.cantilever2d(*(dims or [1000, 1000]), opt_kind=problems.OPT_MMA, filter_kind=problems.FILTER_DENSITY)
    elif kind == "heat":
        P = problems.heat2d(*(dims or [1024, 1024]))
    else:
        P = problems.cantilever3d(*(dims or [96, 48, 48]))
    out = {"problem": P.name, "nelem": P.nelem, "nnode": P.nnode, "setup_host_s": time.time() - t0}
    ctx = capi.Context(0)
    out["device"] = ctx.device_info()
    t0 = time.time()
    S = capi.Simp(ctx, P)
    ctx.sync()
    out["setup_device_s"] = time.time() - t0
    A = S.A
    out["rows"], out["nnz"] = A.rows, A.nnz
    # assembly rate
    rho = ctx.array(np.full(P.nelem, 0.5))
    prm = (P.E0, P.E1, P.poisson, P.penal, P.thickness)
    for _ in range(2):
        A.assemble(S.mesh, S.dofmap, P.eq, prm, P.loads, rho=rho)
    ctx.sync()
    ctx.timer_start()
    for _ in range(5):
        A.assemble(S.mesh, S.dofmap, P.eq, prm, P.loads, rho=rho)
    ms = ctx.timer_stop() / 5
    out["assemble_ms"] = ms
    out["assemble_elems_per_s"] = P.nelem / (ms * 1e-3)
    # SpMV sweep
    bytes_spmv = 12 * A.nnz + 24 * A.rows
    sweep = {}
    for v in (0, 2, 3, 11, 12, 13, 21, 22, 23, 24, 25, 26):
        try:
            ms_f = A.spmv_bench(v, reps=10, flush_l2=True)
            ms_n = A.spmv_bench(v, reps=20, flush_l2=False)
            sweep[str(v)] = {"ms_flushed": ms_f, "gbs_flushed": bytes_spmv / ms_f / 1e6, "ms_b2b": ms_n, "gbs_b2b": bytes_spmv / ms_n / 1e6}
        except capi.Pf2Error as e:
            sweep[str(v)] = {"error": str(e)[:80]}
    out["spmv_bytes"] = bytes_spmv
    out["spmv"] = sweep
    # one solve
    x = ctx.empty(A.rows)
    ctx.sync()
    t0 = time.time()
    ctx.timer_start()
    it, relres = A.solve(capi.SOLVER_SCALINGCG, A.device_F(), x)
    ms = ctx.timer_stop()
    out["solve"] = {"iters": it, "relres": relres, "ms": ms, "ms_per_iter": ms / max(it, 1), "wall_s": time.time() - t0,
                    "pcg_bytes_per_iter": 12 * A.nnz + 112 * A.rows, "gbs": (12 * A.nnz + 112 * A.rows) * it / ms / 1e6}
    # design iterations
    its = []
    for k in range(iters):
        t0 = time.time()
        st = S.iterate(check_convergence=False)
        st["wall_s"] = time.time() - t0
        st["phase_ms"] = S.phase_ms()
        its.append(st)
    out["design_iterations"] = its
    out["launches"] = ctx.launch_count()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", f"probe_{P.name}.json")
    json.dump(out, open(path, "w"), indent=1, default=float)
    print(json.dumps(out, default=float))


if __name__ == "__main__":
    main()
