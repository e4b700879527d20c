"""Tuning probe for the PCG loop: a few design iterations of a bench workload, prints one JSON line with the time per CG iteration
(solve phase / iterations) and the persistent kernel's phase split.   python tools/pcg_tune.py 2m [iterations]
Environment: PF2_LIB (tuning build), PF2_PCG=0 (three-kernel loop), PF2_PCG_GRID, PF2_PCG_CS_MB."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pansfem2_b200 import capi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "2m"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = capi.Context(0)
P = bench.make_problem(name)
S = capi.Simp(ctx, P)
S.iterate(check_convergence=False)
S.A.solver_stats(reset=True)
its, solve_ms = 0, 0.0
for _ in range(n):
    st = S.iterate(check_convergence=False)
    its += st["cg_iters"]; solve_ms += S.phase_ms()["solve"]
pcg, ks = S.A.pcg_stats(), S.A.solver_stats()
print(json.dumps({"workload": name, "lib": os.path.basename(os.environ.get("PF2_LIB", "default")), "pcg": os.environ.get("PF2_PCG", "1"),
                  "grid_env": os.environ.get("PF2_PCG_GRID"), "cg_iters": its, "ms_per_cg_iter": solve_ms / its, "grid": pcg["grid"],
                  "product_ms": round(ks["spmv_ms"], 5), "update_ms": round(ks["update_ms"], 5), "pupdate_ms": round(ks["pupdate_ms"], 5),
                  "wait_ms": [round(pcg["product_wait_ms"], 5), round(pcg["update_wait_ms"], 5), round(pcg["pupdate_wait_ms"], 5)], "f": st["f"]}))
if pcg["solves"] and os.environ.get("PF2_PCG_DBG"):
    import ctypes as C
    import numpy as np
    buf = np.zeros((6, 2048), np.uint64)
    capi._ck(capi.lib().pf2_csr_pcg_debug(S.A.h, buf.ctypes.data_as(C.c_void_p)))
    G = pcg["grid"]
    t = buf[:, :G].astype(np.int64)
    t0 = t[0].min()
    names = ["product", "update", "pupdate"]
    for ph in range(3):
        st_, en = t[2 * ph] - t0, t[2 * ph + 1] - t0
        q = lambda x: [int(v) for v in np.percentile(x, [0, 10, 50, 90, 99, 100])]
        dur = en - st_
        print(json.dumps({"phase": names[ph], "start_ns_pct[0,10,50,90,99,100]": q(st_), "end_ns": q(en), "dur_ns": q(dur), "cta0": [int(st_[0]), int(en[0])],
                          "slowest_ctas": [int(i) for i in np.argsort(en)[-6:]], "fastest_ctas": [int(i) for i in np.argsort(en)[:6]]}))
S.close()
