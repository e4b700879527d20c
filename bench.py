#!/usr/bin/env python
"""bench.py -- SIMP design iterations/s on the B200 hot path, next to the reference CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|2m|c1|c3|c4s] [--impl reference]

One "step" = one design iteration of sample/optimize/sample_optimize_density_mma.cpp's loop (filter -> batched element
stiffness -> CSR assembly -> ScalingCG -> compliance/sensitivities -> filtered sensitivities -> MMA update) on a
synthetic structured mesh, starting from the uniform design s = 0.5.  The default workload is BASELINE.json configs[1]:
2-D plane-strain cantilever 2000x1000 Q4 (4.0 M dof), MMA + density filter, one B200.

Prints ONE JSON line (rank 0).  value = device-resident loop; e2e = the same loop through the host-buffer C-ABI entry
point (design uploaded from / downloaded to pinned host memory every step); roofline = the SpMV(+p.Ap) kernel of the
PCG loop, timed live with CUDA events; cpu_baseline = the unmodified reference (oracle/_ref) on the box's host cores on
a bounded sample, scaled to one design iteration.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, dims, optimiser, filter, description)
    "c2": ("2d", (2000, 1000), "mma", "density", "configs[1]: 2D plane-strain SIMP cantilever 2000x1000 Q4 (4.0M dof), MMA + density filter"),
    "2m": ("2d", (1000, 1000), "mma", "density", "2M-dof headline: 2D plane-strain SIMP cantilever 1000x1000 Q4 (2.0M dof), MMA + density filter"),
    "c1": ("2d", (60, 40), "oc", "heaviside", "configs[0]: sample/optimize cantilever 60x40 Q4, OC + Heaviside filter"),
    "c3": ("heat", (2048, 2048), "oc", "density", "configs[2]: 2D heat-transfer TO 2048x2048 Q4 (4.2M dof), scaled-CG, OC"),
    "c4s": ("3d", (128, 64, 64), "oc", "density", "configs[3] scaled: 3D hex8 cantilever 128x64x64 (1.6M dof)"),
    "c4": ("3d", (256, 128, 128), "oc", "density", "configs[3]: 3D hex8 cantilever 256x128x128 (12.8M dof), OC + density filter"),
    "c5": ("3d", (384, 192, 192), "oc", "density", "configs[4]: 3D hex8 SIMP cantilever 384x192x192 (43M dof), OC + density filter"),
}


def make_problem(name, xr=None):
    """xr = (i0, i1): only that slab of element planes (multi-GPU ranks never build the whole mesh)."""
    from pansfem2_b200 import problems
    kind, dims, opt, flt, _ = WORKLOADS[name]
    opt_kind = problems.OPT_MMA if opt == "mma" else problems.OPT_OC
    fk = problems.FILTER_DENSITY if flt == "density" else problems.FILTER_HEAVISIDE
    if kind == "2d":
        return problems.cantilever2d(*dims, opt_kind=opt_kind, filter_kind=fk, xr=xr)
    if kind == "heat":
        return problems.heat2d(*dims, opt_kind=opt_kind, filter_kind=fk, xr=xr)
    return problems.cantilever3d(*dims, opt_kind=opt_kind, filter_kind=fk, xr=xr)


def global_sizes(name):
    kind, dims, *_ = WORKLOADS[name]
    ndof = {"2d": 2, "heat": 1, "3d": 3}[kind]
    nelem = int(np.prod(dims))
    nnode = int(np.prod([d + 1 for d in dims]))
    return dims, ndof, nelem, nnode


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md's clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return d.get(workload)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(P, cg_iters_per_step, opt_steps_hint=25, budget_s=25.0):
    """Time the reference's own CPU implementation (oracle/_ref, unmodified headers; else the C port) on a bounded sample
    of workload P and scale it to one design iteration:
        t_iter = 3*t_element_pass + t_assembling + t_tocsr + N_cg * t_cg_iteration + t_filter_update
    Element passes and assembly are timed on a strip of the mesh (per-element cost does not depend on mesh size), the CG
    iteration on a matrix of the strip's size class is NOT representative, so it is timed on the largest strip that fits the
    budget and scaled by rows (the reference's CG iteration is bandwidth-bound streaming: time ~ nnz)."""
    from pansfem2_b200 import problems
    from oracle import reflib, portlib
    use_ref = reflib.available()
    kind = "reference" if use_ref else "port"
    cores = os.cpu_count() or 1
    t_begin = time.time()
    # strip: same problem family at reduced size (~1.2 M elements: 10-30 s of reference CPU work in total)
    nel_target = 1200000
    scale = (P.nelem / nel_target) ** (1.0 / len(P.grid))
    dims = [max(4, int(round(g / scale / 2)) * 2) for g in P.grid]
    if P.eq == problems.EQ_PLANESTRAIN:
        Ps = problems.cantilever2d(*dims, opt_kind=P.opt_kind, filter_kind=P.filter_kind)
    elif P.eq == problems.EQ_HEAT:
        Ps = problems.heat2d(*dims, opt_kind=P.opt_kind, filter_kind=P.filter_kind)
    else:
        Ps = problems.cantilever3d(*dims, opt_kind=P.opt_kind, filter_kind=P.filter_kind)
    Emod = np.full(Ps.nelem, P.E1 * 0.5 ** P.penal + P.E0 * (1 - 0.5 ** P.penal))
    if use_ref:
        reflib.set_num_threads(cores)
        S = reflib.assemble(Ps.eq, Ps.coords, Ps.conn, Ps.fixed, Ps.loads, Emod, P.poisson, P.thickness)
        t_el, t_as, t_csr = S.times["element"], S.times["assembling"], S.times["tocsr"]
        F = S.arrays()[3]
        n_it = 30
        # best thread count for SpMV at this size (the reference forks an OpenMP team per product, CSR.h:114)
        best = None
        for thr in sorted({1, min(8, cores), cores}):
            reflib.set_num_threads(thr)
            _, sec, _ = S.solve(1, F, itrmax=n_it)
            if best is None or sec < best[0]:
                best = (sec, thr)
        t_cg, threads = best[0] / n_it, best[1]
        rows_s, nnz_s = S.rows, S.nnz
    else:
        portlib.set_num_threads(cores)
        S, n2g, ufix, tm = portlib.assemble(Ps.eq, Ps.coords, Ps.conn, Ps.fixed, Ps.loads, Emod, P.poisson, P.thickness)
        t_el, t_as, t_csr = tm["element"], tm["scatter"], 0.0
        F = S.arrays()[3]
        t0 = time.time(); S.solve(1, F, itrmax=30); t_cg = (time.time() - t0) / 30
        threads, rows_s, nnz_s = cores, S.rows, S.nnz
    per_elem = (3 * t_el + t_as + t_csr) / Ps.nelem
    # optimiser + filters on the strip
    t0 = time.time()
    if use_ref:
        flt = reflib.Filter(Ps.filter_kind, *Ps.nbrs)
        s = np.full(Ps.nelem, 0.5)
        rho = flt.apply(1.0, s)
        d1 = flt.sens(1.0, s, -np.ones(Ps.nelem)); d2 = flt.sens(1.0, s, np.full(Ps.nelem, 1.0 / (0.5 * Ps.nelem)))
        if Ps.opt_kind == problems.OPT_MMA:
            m = reflib.MMA(Ps.nelem, 1, P.mma[7], [P.mma[8]], [P.mma[9]], [P.mma[10]], P.mma[11], P.mma[12])
            m.set_parameters(*P.mma[:7])
            m.update(s, 1.0, d1, [rho.sum() / (0.5 * Ps.nelem) - 1.0], d2[None, :])
        else:
            o = reflib.OC(Ps.nelem, *P.oc)
            o.update(flt, 1.0, P.weightlimit, P.scale1, s, 1.0, d1, 0.0, d2)
    t_opt_per_elem = (time.time() - t0) / Ps.nelem
    nnz_full = P.extra.get("nnz")
    t_iter = per_elem * P.nelem + t_opt_per_elem * P.nelem + cg_iters_per_step * t_cg * (nnz_full / nnz_s if nnz_full else P.nelem / Ps.nelem)
    sample = (f"{'unmodified reference headers (oracle/_ref)' if use_ref else 'C port (oracle/pf2_oracle.c)'}: element+Assembling+CSR, "
              f"filters+optimiser and 30 ScalingCG iterations timed on a {'x'.join(map(str, dims))} strip of the same problem "
              f"({Ps.nelem} elements, {rows_s} dof), scaled per element / per nonzero to the full mesh with the GPU run's "
              f"{cg_iters_per_step:.0f} CG iterations per design iteration; sample took {time.time() - t_begin:.1f} s")
    detail = {"element_us": 1e6 * t_el / Ps.nelem, "assembling_us": 1e6 * t_as / Ps.nelem, "tocsr_us": 1e6 * t_csr / Ps.nelem,
              "cg_iter_ms_strip": 1e3 * t_cg, "spmv_threads": threads, "strip_rows": rows_s, "strip_nnz": nnz_s}
    return {"value": 1.0 / t_iter, "unit": "design iterations/s", "cores": threads, "host_cores": cores, "kind": kind, "sample": sample,
            "detail": detail}


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-matrix-free-leg", action="store_true", help="skip the extra timed leg with the matrix-free operator")
    ap.add_argument("--operator", default=os.environ.get("PF2_OPERATOR", "csr"), choices=["csr", "matrix-free"],
                    help="how K is applied inside the PCG: the assembled CSR (default; the path the roofline is quoted on) or the opt-in "
                         "matrix-free operator for uniform structured meshes (pf2_csr_matrix_free)")
    ap.add_argument("--cg-iters-hint", type=float, default=0.0, help="CG iterations per design iteration for --impl reference scaling")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    desc = WORKLOADS[args.workload][4]
    base = {"metric": "SIMP design iterations/s", "unit": "design iterations/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic structured mesh, uniform initial design s=0.5"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        P = make_problem(args.workload)
        P.extra["nnz"] = estimate_nnz(P)
        hint = args.cg_iters_hint or estimate_cg_iters(P)
        t0 = time.time()
        vals = []
        for _ in range(max(1, min(args.steps, 2))):
            cb = cpu_reference_sample(P, hint)
            vals.append(cb["value"])
        v = float(np.mean(vals))
        cb["value"] = v
        out = dict(base, impl="reference", value=v, ms_per_step=1e3 / v,
                   config={"workload": desc, "elements": P.nelem, "dof": P.free_dofs(), "parallelism": "host CPU",
                           "cg_iters_per_step_assumed": hint},
                   cpu_baseline=cb, e2e={"value": v, "unit": base["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                   gpu_launches=0, wall_s=time.time() - t0)
        print(json.dumps(out))
        return 0

    import torch
    import torch.distributed as dist
    from pansfem2_b200 import capi

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    ctx = capi.Context(local_rank)
    dims, ndof, nelem_global, nnode_global = global_sizes(args.workload)
    if world > 1:
        # one problem, row-block (x-slab) partitioned over the ranks: strong scaling
        from pansfem2_b200 import partition
        D = capi.Dist(ctx, rank, world)
        slab = partition.slab_from_factory(lambda xr: make_problem(args.workload, xr=xr), dims, ndof, rank, world)
        P = slab.local
        S = capi.Simp(ctx, P, matrix_free=(args.operator == "matrix-free"))
        D.set_simp_partition(S, slab, nelem_global)
    else:
        P = make_problem(args.workload)
        S = capi.Simp(ctx, P, matrix_free=(args.operator == "matrix-free"))
    P.extra["nnz"] = S.A.nnz
    nelem = P.nelem
    s_in, s_out, rho_out = capi.pinned_empty(nelem), capi.pinned_empty(nelem), capi.pinned_empty(nelem)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    hist = []
    for _ in range(args.warmup):
        hist.append(S.iterate(check_convergence=False))
    S.A.solver_stats(reset=True)
    sampler = ClockSampler(local_rank)
    # ---- timed region 1: device-resident loop ----
    barrier()
    sampler.start()
    l0 = ctx.launch_count()
    ctx.timer_start()
    t_wall = time.time()
    steps = []
    for _ in range(args.steps):
        st = S.iterate(check_convergence=False)
        st["phase_ms"] = S.phase_ms()
        steps.append(st)
    ms = ctx.timer_stop()
    barrier()
    wall = time.time() - t_wall
    launches = ctx.launch_count() - l0
    kstats = S.A.solver_stats(reset=True)
    # ---- timed region 2: end to end through host buffers ----
    out_state = S.get()
    s_in[:] = out_state["s"]
    barrier()
    ctx.timer_start()
    e2e_steps = []
    for _ in range(args.steps):
        st = S.iterate_host(s_in, s_out, rho_out, check_convergence=False)
        e2e_steps.append(st)
        s_in[:] = s_out
    ms_e2e = ctx.timer_stop()
    barrier()
    clocks = sampler.stop()
    # ---- extra leg (reported beside the headline, never instead of it): the same loop with K applied matrix-free ----
    mf = None
    if args.operator == "csr" and not args.no_matrix_free_leg:
        try:
            S.A.matrix_free(S.mesh, S.dofmap, P.eq)
            S.iterate(check_convergence=False)                       # first assembly with the operator (Ke0 upload), untimed
            S.A.solver_stats(reset=True)
            barrier()
            ctx.timer_start()
            mf_steps = [S.iterate(check_convergence=False) for _ in range(args.steps)]
            ms_mf = ctx.timer_stop()
            barrier()
            ks = S.A.solver_stats(reset=True)
            if world > 1:
                t = torch.tensor([ms_mf], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_mf = t.item()
            mf = {"value": args.steps / (ms_mf * 1e-3), "unit": base["unit"], "ms_per_step": ms_mf / args.steps,
                  "cg_iters_per_step": float(np.mean([s["cg_iters"] for s in mf_steps])), "cg_relres_max": max(s["cg_relres"] for s in mf_steps),
                  "operator_ms": ks["spmv_ms"], "update_ms": ks["update_ms"], "pupdate_ms": ks["pupdate_ms"],
                  "note": "opt-in pf2_csr_matrix_free (uniform structured mesh): y = sum_e E_e Ke0 p_e instead of the CSR stream; same assembly, "
                          "preconditioner, recurrences and stopping test; parity in tests/test_gpu_matrix_free.py"}
        except capi.Pf2Error as e:
            mf = {"unavailable": str(e)[:160]}
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    value = args.steps / (ms * 1e-3)               # one (possibly partitioned) problem: whole-job design iterations/s
    e2e_value = args.steps / (ms_e2e * 1e-3)
    cg_iters = float(np.mean([s["cg_iters"] for s in steps]))
    peak, peak_src = measured_peak()
    spmv_bytes = 12 * S.A.nnz + 24 * S.A.rows
    roof = None
    if world > 1:
        tt = torch.tensor([float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt)
        launches = int(tt.item())
    if kstats["samples"] > 0 and kstats["spmv_ms"] > 0:
        achieved = spmv_bytes / (kstats["spmv_ms"] * 1e-3) / 1e9
        v = kstats["variant"]
        kname = "spmv_sell_kernel<DOT> (SELL-32, thread per row)" if v == 31 else ("spmv_tma_kernel" if v >= 21 else "spmv_stream_kernel" if v >= 11 else "spmv_vector_kernel")
        if v == 41:
            # matrix-free operator: the matrix stream is gone, the dominant HBM kernel of a PCG iteration is the fused vector update
            upd_bytes = 64 * S.A.rows
            achieved = upd_bytes / (kstats["update_ms"] * 1e-3) / 1e9
            mf_bytes = (8 + 8 + 4) * S.A.rows + 8 * nelem
            roof = {"bound": "hbm", "kernel": "cg_update_kernel<Jacobi>: x += a p, r -= a Kp, z = r/D, z.r, r.r (the largest HBM kernel once K is applied matrix-free)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": None,
                    "algorithmic_bytes_per_launch": upd_bytes, "avg_launch_ms": kstats["update_ms"], "samples": kstats["samples"],
                    "matrix_free_operator": {"kernel": "spmv_mf_kernel (variant 41): y = sum_e E_e Ke0 p_e fused with p.Kp, DFMA-bound",
                                             "avg_launch_ms": kstats["spmv_ms"], "compulsory_bytes_per_launch": mf_bytes,
                                             "dfma_gflops": 2.0 * S.A.nnz / (kstats["spmv_ms"] * 1e-3) / 1e9,
                                             "csr_equivalent_GBps": spmv_bytes / (kstats["spmv_ms"] * 1e-3) / 1e9},
                    "pcg_iteration": {"ms": kstats["spmv_ms"] + kstats["update_ms"] + kstats["pupdate_ms"], "update_ms": kstats["update_ms"],
                                      "pupdate_ms": kstats["pupdate_ms"]}}
        else:
            roof = None
    if roof is None and kstats["samples"] > 0 and kstats["spmv_ms"] > 0:
        achieved = spmv_bytes / (kstats["spmv_ms"] * 1e-3) / 1e9
        v = kstats["variant"]
        kname = "spmv_sell_kernel<DOT> (SELL-32, thread per row)" if v == 31 else ("spmv_tma_kernel" if v >= 21 else "spmv_stream_kernel" if v >= 11 else "spmv_vector_kernel")
        roof = {"bound": "hbm", "kernel": f"{kname}: y = K p fused with p.Kp (variant {v})", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": ncu_traffic(args.workload),
                "algorithmic_bytes_per_launch": spmv_bytes, "avg_launch_ms": kstats["spmv_ms"], "samples": kstats["samples"],
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "pcg_iteration": {"bytes": 12 * S.A.nnz + 112 * S.A.rows, "ms": kstats["spmv_ms"] + kstats["update_ms"] + kstats["pupdate_ms"],
                                  "update_ms": kstats["update_ms"], "pupdate_ms": kstats["pupdate_ms"]}}
    out = dict(base, value=value, ms_per_step=ms / args.steps,
               config={"workload": desc, "elements": nelem_global, "dof_local": S.A.rows, "nnz_local": S.A.nnz,
                       "parallelism": "1 GPU" if world == 1 else (f"{world} x-slabs (row-block partition; PCG halo + allreduce fused into the kernels over NVLink peer memory, "
                                                                    "NCCL for the per-design-iteration exchanges)" if os.environ.get("PF2_P2P", "1") != "0"
                                                                    else f"{world} x-slabs (row-block partition, NCCL halo exchange + allreduce)"),
                       "l2": "working set (CSR values+indices %.0f MB) exceeds the 126 MB L2; no flush needed" % (12 * S.A.nnz / 1e6),
                       "cg_iters_per_step": cg_iters, "solver": "ScalingCG eps=1e-10 x0=0",
                       "operator": "assembled CSR (SELL-32 mirror)" if args.operator == "csr" else "matrix-free on the uniform mesh (pf2_csr_matrix_free); CSR still assembled every iteration"},
               e2e={"value": e2e_value, "unit": base["unit"], "h2d_bytes_per_step": int(8 * nelem), "d2h_bytes_per_step": int(16 * nelem),
                    "ms_per_step": ms_e2e / args.steps},
               gpu_launches=int(launches), clocks=clocks, roofline=roof, wall_s=wall,
               phases_ms=steps[-1]["phase_ms"], objective=[s["f"] for s in steps], cg_relres_max=max(s["cg_relres"] for s in steps))
    # third figure of BASELINE.json's metric: assembly elements/s (numeric phase: element routine + scatter / gather into the CSR),
    # from the per-phase device time of the timed steps (this rank's elements; slabs assemble concurrently)
    asm_ms = float(np.mean([st["phase_ms"]["assemble"] for st in steps]))
    if asm_ms > 0:
        bytes_per_elem = {2: 590.0, 1: 175.0, 3: 4300.0}.get(P.ndof, 0.0)     # Q4 plane strain / heat / hex8, map included
        out["assembly"] = {"elements_per_s": nelem / (asm_ms * 1e-3), "ms": asm_ms, "elements_local": int(nelem),
                           "algorithmic_bytes_per_element": bytes_per_elem, "algorithmic_GBps": bytes_per_elem * nelem / (asm_ms * 1e-3) / 1e9,
                           "kernel": "assemble_gather_kernel (row gather, every entry written once, bitwise reproducible)"
                                     if (P.ndof < 3 and not os.environ.get("PF2_ASSEMBLE_SCATTER")) else "assemble_kernel (scatter, fp64 RED)"}
    if mf is not None:
        out["matrix_free"] = mf
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            try:
                out["cpu_baseline"] = cpu_reference_sample(P, cg_iters)
            except Exception as e:  # the checker must never take the bench down
                out["cpu_baseline"] = {"error": repr(e)[:200]}
        print(json.dumps(out, default=float))
    S.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def estimate_nnz(P):
    ndof = P.ndof
    pairs = 1
    for g in P.grid:
        pairs *= 3 * (g + 1) - 2
    return pairs * ndof * ndof


def estimate_cg_iters(P):
    """Jacobi-PCG iterations to 1e-10 on the uniform design grow ~ linearly with the longest mesh edge (measured: 357 @ 60x40,
    2135 @ 400x200, 6693 @ 1000x1000)."""
    return 6.7 * max(P.grid)


if __name__ == "__main__":
    sys.exit(main())
