#!/usr/bin/env python
"""bench.py -- SIMP design iterations/s on the B200 hot path, next to the reference CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|2m|c1|c3|c4s|c4|c5] [--impl reference]
                    [--no-extra-legs] [--no-headline-2m] [--no-hex8] [--no-cpu-baseline]

One "step" = one design iteration of sample/optimize/sample_optimize_density_mma.cpp:83-208 (filter -> batched element
stiffness -> CSR assembly -> ScalingCG -> compliance/sensitivities -> filtered sensitivities -> MMA update) on a synthetic
structured mesh, starting from the uniform design s = 0.5.  The headline workload is BASELINE.json configs[1]: 2-D plane-strain
cantilever 2000x1000 Q4 (4.0 M dof), MMA + density filter.

Prints ONE JSON line (rank 0).  Every timed leg starts from the SAME state: the design, the Heaviside beta and the optimiser are
reset (pf2_simp_reset), W warm-up iterations run untimed, then K iterations are timed on the device (CUDA events on the library's
stream, max over ranks) -- so `value`, `e2e` and the extra legs time the same design iterations k = W .. W+K-1.

    value          device-resident loop, the reference's algorithm as is (assembled CSR, ScalingCG from x0 = 0)
    e2e            the same iterations through pf2_simp_iterate_host (design uploaded from / downloaded to pinned host memory)
    legs           opt-in variants, each with its own cg_iters_per_step: warm_start (pf2_solve_x0), matrix_free
                   (pf2_csr_matrix_free), warm_start+matrix_free; at N > 1 also single_reduction_cg (pf2_csr_set_cg_variant) and all three together
    headline_2m    BASELINE.json's metric size: 1000x1000 Q4 (2.0 M dof), value + e2e
    hex8_scaling   configs[3] 256x128x128 hex8 (one SIMP iteration from the uniform design) and configs[4] 384x192x192 hex8
                   (design loop), row-partitioned over the N GPUs like everything else at N > 1
    roofline       the dominant kernel: the persistent PCG kernel (one launch per solve), algorithmic bytes per launch =
                   CG iterations x (12 nnz + 108 rows) over its CUDA-event duration; `spmv` inside it = the product
                   phase timed by the kernel's own %globaltimer stamps and the product kernel timed alone with CUDA events
    cpu_baseline   the unmodified reference (oracle/_ref) on the box's host cores: a bounded strip scaled to one design iteration
                   ("extrapolated": true) and one REAL design iteration at 400x200 next to the same iteration on the GPU
"""
from __future__ import annotations

import argparse
import gc
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
    "pair": ("2d", (400, 200), "mma", "density", "measured pair: 2D plane-strain SIMP cantilever 400x200 Q4 (160k dof), MMA + density filter"),
}
UNIT = "design iterations/s"
OBJECTIVE_N1 = os.path.join(ROOT, "tests", "golden", "bench_objective_n1.json")
MEASURED_CG = os.path.join(ROOT, "profiles", "measured_cg_iters.json")


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


def ncu_traffic(key):
    """dram bytes of the dominant kernel from the committed ncu --set full capture (profiles/ncu_traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(key)
    except Exception:
        return None


def load_json(path, default):
    try:
        return json.load(open(path))
    except Exception:
        return default


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline
# ---------------------------------------------------------------------------------------------------------------------
def strip_problem(P, nel_target):
    """The same problem family at reduced size (same aspect ratio), about nel_target elements."""
    from pansfem2_b200 import problems
    scale = max(1.0, (P.nelem / nel_target)) ** (1.0 / len(P.grid))
    dims = [max(4, int(round(g / scale / 2)) * 2) for g in P.grid]
    if P.eq == problems.EQ_PLANESTRAIN:
        Ps = problems.cantilever2d(*dims, opt_kind=P.opt_kind, filter_kind=P.filter_kind)
    elif P.eq == problems.EQ_HEAT:
        Ps = problems.heat2d(*dims, opt_kind=P.opt_kind, filter_kind=P.filter_kind)
    else:
        Ps = problems.cantilever3d(*dims, opt_kind=P.opt_kind, filter_kind=P.filter_kind)
    return Ps, dims


def cpu_reference_sample(P, cg_iters_per_step, cg_source, nel_target=400000, cg_its=20):
    """Time the reference's own CPU implementation (oracle/_ref, unmodified headers; else the C port) on a bounded strip of workload P
    and scale it to one design iteration of the full mesh (EXTRAPOLATED: a full-size reference iteration takes ~10 minutes):
        t_iter = (3 t_element + t_assembling + t_tocsr) nelem  +  t_filters+optimiser nelem  +  N_cg t_cg_iteration nnz / nnz_strip
    Per-element costs do not depend on the mesh size; the reference's CG iteration is a bandwidth-bound stream (time ~ nnz)."""
    from pansfem2_b200 import problems
    from oracle import reflib, portlib
    use_ref = reflib.available()
    kind = "reference" if use_ref else "port"
    cores = os.cpu_count() or 1
    t_begin = time.time()
    Ps, dims = strip_problem(P, nel_target)
    Emod = np.full(Ps.nelem, P.E1 * 0.5 ** P.penal + P.E0 * (1 - 0.5 ** P.penal))
    if use_ref:
        reflib.set_num_threads(cores)
        S = reflib.assemble(Ps.eq, Ps.coords, Ps.conn, Ps.fixed, Ps.loads, Emod, P.poisson, P.thickness)
        t_el, t_as, t_csr = S.times["element"], S.times["assembling"], S.times["tocsr"]
        F = S.arrays()[3]
        # best thread count for SpMV at this size (the reference forks an OpenMP team per product, CSR.h:114)
        best = None
        for thr in sorted({1, min(8, cores), cores}):
            reflib.set_num_threads(thr)
            _, sec, _ = S.solve(1, F, itrmax=cg_its)
            if best is None or sec < best[0]:
                best = (sec, thr)
        t_cg, threads = best[0] / cg_its, best[1]
        rows_s, nnz_s = S.rows, S.nnz
    else:
        portlib.set_num_threads(cores)
        S, n2g, ufix, tm = portlib.assemble(Ps.eq, Ps.coords, Ps.conn, Ps.fixed, Ps.loads, Emod, P.poisson, P.thickness)
        t_el, t_as, t_csr = tm["element"], tm["scatter"], 0.0
        F = S.arrays()[3]
        t0 = time.time(); S.solve(1, F, itrmax=cg_its); t_cg = (time.time() - t0) / cg_its
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
    sample = (f"EXTRAPOLATED from a strip: {'unmodified reference headers (oracle/_ref)' if use_ref else 'C port (oracle/pf2_oracle.c)'}: "
              f"element+Assembling+CSR, filters+optimiser and {cg_its} ScalingCG iterations timed on a {'x'.join(map(str, dims))} mesh of the same "
              f"problem ({Ps.nelem} elements, {rows_s} dof), scaled per element / per nonzero to the full mesh with {cg_iters_per_step:.0f} CG "
              f"iterations per design iteration ({cg_source}); sample took {time.time() - t_begin:.1f} s")
    detail = {"element_us": 1e6 * t_el / Ps.nelem, "assembling_us": 1e6 * t_as / Ps.nelem, "tocsr_us": 1e6 * t_csr / Ps.nelem,
              "cg_iter_ms_strip": 1e3 * t_cg, "spmv_threads": threads, "strip_rows": rows_s, "strip_nnz": nnz_s}
    return {"value": 1.0 / t_iter, "unit": UNIT, "cores": threads, "host_cores": cores, "kind": kind, "sample": sample,
            "extrapolated": True, "strip": "x".join(map(str, dims)), "cg_iters_per_step": cg_iters_per_step, "cg_iters_source": cg_source,
            "detail": detail}


def cpu_reference_real_iteration(name="pair"):
    """ONE real design iteration (k = 0) of the unmodified reference's loop on the `pair` workload: nothing modelled."""
    from oracle import reflib, portlib
    P = make_problem(name)
    use_ref = reflib.available()
    cores = os.cpu_count() or 1
    run = reflib.simp_run if use_ref else portlib.simp_run
    (reflib if use_ref else portlib).set_num_threads(min(8, cores))
    t0 = time.time()
    R = run(P.eq, P.coords, P.conn, P.fixed, P.loads, P.filter_kind, P.nbrs, P.opt_kind, P.optp(), P.params(), 1,
            np.full(P.nelem, P.s0), check_convergence=False)
    sec = time.time() - t0
    return {"workload": WORKLOADS[name][4], "seconds": sec, "value": 1.0 / sec, "unit": UNIT, "threads": min(8, cores),
            "kind": "reference" if use_ref else "port", "objective": float(R["hist"][0, 0]), "extrapolated": False}


def reference_cg_count(workload, warmup, steps, P):
    """CG iterations per design iteration for the reference-arm scaling: the count the native arm MEASURED on this workload for the same
    (warmup, steps) window (profiles/measured_cg_iters.json, written from GPU runs), else the growth law fitted to r01's runs."""
    d = load_json(MEASURED_CG, {})
    ent = d.get(workload)
    if ent:
        per = ent.get("cg_iters_by_k")
        if per and len(per) >= warmup + steps:
            return float(np.mean(per[warmup:warmup + steps])), f"measured by the native arm, design iterations {warmup}..{warmup + steps - 1} ({ent.get('source', 'profiles/measured_cg_iters.json')})"
        if per:
            return float(np.mean(per[min(warmup, len(per) - 1):])), f"measured by the native arm on fewer iterations ({ent.get('source', '')})"
    return 6.7 * max(P.grid), "ESTIMATED: 6.7 x longest mesh edge (357 @ 60x40, 2135 @ 400x200, 6693 @ 1000x1000 measured)"


def estimate_nnz(P):
    pairs = 1
    for g in P.grid:
        pairs *= 3 * (g + 1) - 2
    return pairs * P.ndof * P.ndof


# ---------------------------------------------------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------------------------------------------------
class Runner:
    """One workload on this rank's GPU (the whole mesh, or this rank's x-slab of it)."""

    def __init__(self, env, name, matrix_free=False):
        from pansfem2_b200 import capi
        self.env, self.name = env, name
        self.desc = WORKLOADS[name][4]
        self.dims, self.ndof, self.nelem_global, self.nnode_global = global_sizes(name)
        ctx, world, rank = env["ctx"], env["world"], env["rank"]
        if world > 1:
            from pansfem2_b200 import partition
            slab = partition.slab_from_factory(lambda xr: make_problem(name, xr=xr), self.dims, self.ndof, rank, world)
            self.P = slab.local
            self.S = capi.Simp(ctx, self.P, matrix_free=matrix_free)
            env["D"].set_simp_partition(self.S, slab, self.nelem_global)
        else:
            self.P = make_problem(name)
            self.S = capi.Simp(ctx, self.P, matrix_free=matrix_free)
        self.P.extra["nnz"] = self.S.A.nnz
        self.nelem = self.P.nelem
        self.pinned = None
        self.objective = {}

    def barrier(self):
        import torch
        self.env["ctx"].sync()
        torch.cuda.synchronize()
        if self.env["world"] > 1:
            import torch.distributed as dist
            dist.barrier()

    def max_over_ranks(self, vals):
        if self.env["world"] == 1:
            return list(vals)
        import torch
        import torch.distributed as dist
        t = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def sum_over_ranks(self, v):
        if self.env["world"] == 1:
            return v
        import torch
        import torch.distributed as dist
        t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        return t.item()

    def leg(self, steps, warmup, host=False, warm_start=False, matrix_free=False, repeat_first=False, single_reduction=False, tag="value"):
        """Reset to the uniform design, `warmup` untimed iterations, `steps` timed ones.  repeat_first: every iteration (warm-up and
        timed) is design iteration k = 0 from the uniform design (configs[3]: 'single SIMP iteration'), resets outside the timer."""
        from pansfem2_b200 import capi
        S, ctx = self.S, self.env["ctx"]
        if matrix_free != getattr(self, "_mf", False):
            if matrix_free:
                S.A.matrix_free(S.mesh, S.dofmap, self.P.eq)
            else:
                S.A.set_spmv_variant(0)
            self._mf = matrix_free
        S.set_warm_start(warm_start)
        S.A.set_cg_variant(1 if single_reduction else 0)     # N > 1, peer-memory backend: one cross-GPU sum and two kernels per PCG iteration
        S.reset()
        if host and self.pinned is None:
            self.pinned = (capi.pinned_empty(self.nelem), capi.pinned_empty(self.nelem), capi.pinned_empty(self.nelem))
        if host:
            s_in, s_out, rho_out = self.pinned
            s_in[:] = self.P.s0

        def one():
            if host:
                st = S.iterate_host(s_in, s_out, rho_out, check_convergence=False)
                s_in[:] = s_out
                return st
            return S.iterate(check_convergence=False)

        hist = []
        for _ in range(warmup):
            if repeat_first:
                S.reset()
            hist.append(one())
        S.A.solver_stats(reset=True)
        self.barrier()
        l0 = ctx.launch_count()
        t_wall = time.time()
        steps_out, ms = [], 0.0
        if repeat_first:
            for _ in range(steps):
                S.reset()
                self.barrier()
                ctx.timer_start()
                st = one()
                ms += ctx.timer_stop()
                st["phase_ms"] = S.phase_ms()
                steps_out.append(st)
        else:
            ctx.timer_start()
            for _ in range(steps):
                st = one()
                st["phase_ms"] = S.phase_ms()
                steps_out.append(st)
            ms = ctx.timer_stop()
        self.barrier()
        wall = time.time() - t_wall
        launches = ctx.launch_count() - l0
        ks, pcg = S.A.solver_stats(), S.A.pcg_stats()
        (ms,) = self.max_over_ranks([ms])
        objective = [h["f"] for h in hist] + [s["f"] for s in steps_out]
        self.objective[tag] = objective
        return {"ms": ms, "steps": steps_out, "launches": launches, "kstats": ks, "pcg": pcg, "wall": wall,
                "value": steps / (ms * 1e-3), "ms_per_step": ms / steps,
                "cg_iters_per_step": float(np.mean([s["cg_iters"] for s in steps_out])),
                "cg_iters_by_k": [h["cg_iters"] for h in hist] + [s["cg_iters"] for s in steps_out],
                "cg_relres_max": max(s["cg_relres"] for s in steps_out), "objective": objective}

    def roofline(self, L):
        """The dominant kernel of leg L and the SpMV inside it."""
        from pansfem2_b200 import capi  # noqa: F401
        A, world = self.S.A, self.env["world"]
        peak, peak_src = measured_peak()
        ks, pcg = L["kstats"], L["pcg"]
        rows_own, nnz = A.rows, A.nnz          # local slab (ghost rows included: < 2 planes)
        spmv_bytes = 12 * nnz + 20 * rows_own
        iter_bytes = 12 * nnz + 108 * rows_own
        out = None
        if pcg["solves"] > 0 and pcg["kernel_ms"] > 0:
            launch_ms = pcg["kernel_ms"] / pcg["solves"]
            its_per_launch = pcg["iters"] / pcg["solves"]
            bytes_per_launch = its_per_launch * iter_bytes
            achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
            traffic_it = ncu_traffic("pcg_" + self.name) if world == 1 else None
            out = {"bound": "hbm",
                   "kernel": "pcg_persistent_kernel: one cooperative launch per ScalingCG solve (SELL-32 product + p.Kp, fused vector updates, grid barriers carry the dot products)",
                   "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                   "traffic": (traffic_it["dram_bytes_per_iteration"] * its_per_launch) if traffic_it else None,
                   "frac_physical": (traffic_it["dram_bytes_per_iteration"] * its_per_launch / (launch_ms * 1e-3) / 1e9 / peak) if traffic_it else None,
                   "algorithmic_bytes_per_launch": bytes_per_launch, "algorithmic_bytes_per_cg_iteration": iter_bytes,
                   "cg_iterations_per_launch": its_per_launch, "avg_launch_ms": launch_ms, "launches_timed": pcg["solves"], "grid_ctas": pcg["grid"],
                   "ms_per_cg_iteration": pcg["kernel_ms"] / max(pcg["iters"], 1), "frac_of_nominal_8TBs": achieved / 8000.0,
                   "timing": "CUDA events on the library's stream around every cooperative launch of the timed steps",
                   "loop_overhead": {"solve_phase_ms_per_step": float(np.mean([s["phase_ms"]["solve"] for s in L["steps"]])),
                                     "kernel_ms_per_step": pcg["kernel_ms"] / max(len(L["steps"]), 1)},
                   "spmv": {"in_kernel_phase_ms": pcg["product_ms"], "algorithmic_bytes": spmv_bytes,
                            "achieved_GBps": spmv_bytes / (pcg["product_ms"] * 1e-3) / 1e9 if pcg["product_ms"] > 0 else None,
                            "frac": spmv_bytes / (pcg["product_ms"] * 1e-3) / 1e9 / peak if pcg["product_ms"] > 0 else None,
                            "note": "product phase of the persistent kernel incl. its grid barrier (+ allreduce at N > 1), %globaltimer of CTA 0",
                            "update_phase_ms": pcg["update_ms"], "pupdate_phase_ms": pcg["pupdate_ms"],
                            "exchange_wait_ms_cta0": [pcg["product_wait_ms"], pcg["update_wait_ms"], pcg["pupdate_wait_ms"]]}}
            if world == 1:
                try:
                    ms_alone = A.spmv_bench(0, reps=10, flush_l2=(12 * nnz < 400e6))
                    tr = ncu_traffic(self.name)
                    out["spmv"].update(alone_ms=ms_alone, alone_GBps=spmv_bytes / (ms_alone * 1e-3) / 1e9, alone_frac=spmv_bytes / (ms_alone * 1e-3) / 1e9 / peak,
                                       alone_traffic=tr, alone_frac_physical=(tr / (ms_alone * 1e-3) / 1e9 / peak) if tr else None,
                                       alone_note="spmv_sell_kernel (same slice routine) launched alone, CUDA events, 10 back-to-back launches")
                except Exception as e:  # noqa: BLE001
                    out["spmv"]["alone_error"] = repr(e)[:120]
        elif ks["samples"] > 0 and ks["spmv_ms"] > 0:
            v = ks["variant"]
            if v == 41:
                upd_bytes = 64 * rows_own
                achieved = upd_bytes / (ks["update_ms"] * 1e-3) / 1e9
                out = {"bound": "hbm", "kernel": "cg_update_kernel<Jacobi> (largest HBM kernel once K is applied matrix-free)",
                       "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": None,
                       "algorithmic_bytes_per_launch": upd_bytes, "avg_launch_ms": ks["update_ms"], "samples": ks["samples"],
                       "operator_ms": ks["spmv_ms"], "csr_equivalent_GBps": spmv_bytes / (ks["spmv_ms"] * 1e-3) / 1e9}
            else:
                achieved = spmv_bytes / (ks["spmv_ms"] * 1e-3) / 1e9
                tr = ncu_traffic(self.name) if world == 1 else None
                out = {"bound": "hbm", "kernel": f"SpMV variant {v} fused with p.Kp (three-kernel PCG loop)", "achieved": achieved, "peak": peak,
                       "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": tr,
                       "frac_physical": (tr / (ks["spmv_ms"] * 1e-3) / 1e9 / peak) if tr else None,
                       "algorithmic_bytes_per_launch": spmv_bytes, "avg_launch_ms": ks["spmv_ms"], "samples": ks["samples"],
                       "pcg_iteration": {"ms": ks["spmv_ms"] + ks["update_ms"] + ks["pupdate_ms"], "update_ms": ks["update_ms"], "pupdate_ms": ks["pupdate_ms"]}}
        return out

    def ilu_probe(self, ilu_itrmax=None):
        """ScalingCG against ILU0CG (CG.h:320-352) on the matrix of the last assembly: iterations and device time of one solve each.
        ilu_itrmax caps the ILU solve (large meshes: thousands of dependency levels per sweep) - then only its ms per iteration counts."""
        from pansfem2_b200 import capi
        ctx, A = self.env["ctx"], self.S.A
        x = ctx.empty(A.rows)
        out = {}
        for tag, solver, cap in (("scalingcg", capi.SOLVER_SCALINGCG, None), ("ilu0cg", capi.SOLVER_ILU0CG, ilu_itrmax)):
            if solver == capi.SOLVER_ILU0CG:
                ctx.timer_start()
                capi._ck(capi.lib().pf2_ilu0_factor(A.h))
                out["ilu0_factor_ms"] = ctx.timer_stop()
            ctx.timer_start()
            try:
                it, rr = A.solve(solver, A.device_F(), x, itrmax=cap or 100000)
                conv = True
            except capi.Pf2Error as e:
                if e.code != capi.E_NOCONV:
                    raise
                it, rr, conv = cap, None, False
            ms = ctx.timer_stop()
            out[tag] = {"iterations": it, "ms": ms, "ms_per_iteration": ms / max(it, 1), "converged": conv, "relres": rr}
        x.free()
        if out["ilu0cg"]["converged"]:
            out["ilu_vs_jacobi_time"] = out["ilu0cg"]["ms"] / out["scalingcg"]["ms"]
        out["note"] = ("ILU0CG = factorisation once (one launch per dependency level) + two triangular sweeps per iteration, each ONE launch in which a CTA "
                       "(or an 8-CTA cluster) walks the levels; natural ordering gives ~(nx+ny) levels, so a sweep is latency-bound: a parity feature of "
                       "the reference, not the fast path")
        return out

    def parity_vs_n1(self, tag="value"):
        ref = load_json(OBJECTIVE_N1, {}).get(self.name)
        mine = self.objective.get(tag)
        if not ref or not mine:
            return None
        n = min(len(ref), len(mine))
        d = max(abs(a - b) / abs(b) for a, b in zip(mine[:n], ref[:n]))
        return {"max_rel_diff": d, "iterations_compared": n, "ok": bool(d < 1e-8), "reference": "tests/golden/bench_objective_n1.json (1 GPU)"}

    def close(self):
        if self.env["world"] > 1:
            self.env["D"].release_p2p(self.S.A)       # peers unmap this matrix' slab before anybody frees it
        self.S.close()
        self.S = None
        self.pinned = None
        gc.collect()


def brief(L, extra=None):
    d = {"value": L["value"], "unit": UNIT, "ms_per_step": L["ms_per_step"], "steps": len(L["steps"]), "cg_iters_per_step": L["cg_iters_per_step"],
         "cg_relres_max": L["cg_relres_max"]}
    if extra:
        d.update(extra)
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the warm-start / matrix-free legs")
    ap.add_argument("--no-headline-2m", action="store_true", help="skip the 1000x1000 leg")
    ap.add_argument("--no-hex8", action="store_true", help="skip the configs[3] / configs[4] hex8 legs")
    ap.add_argument("--hex8", default="c4,c5", help="which hex8 legs to run (comma separated: c4s, c4, c5)")
    ap.add_argument("--cg-iters-hint", type=float, default=0.0, help="CG iterations per design iteration for --impl reference scaling")
    ap.add_argument("--record", default="", help="write the objective / CG-iteration histories of this run to this JSON file (N = 1 runs feed "
                                                 "tests/golden/bench_objective_n1.json and profiles/measured_cg_iters.json)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    desc = WORKLOADS[args.workload][4]
    base = {"metric": "SIMP design iterations/s", "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic structured mesh, uniform initial design s=0.5"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        P = make_problem(args.workload)
        P.extra["nnz"] = estimate_nnz(P)
        if args.cg_iters_hint:
            hint, src = args.cg_iters_hint, "--cg-iters-hint"
        else:
            hint, src = reference_cg_count(args.workload, args.warmup, args.steps, P)
        t0 = time.time()
        # every step is one bounded sample (~5 s of host work); cap the whole arm at ~4 minutes
        vals, cb = [], None
        nsamples = max(1, args.warmup + args.steps)
        for i in range(nsamples):
            cb = cpu_reference_sample(P, hint, src)
            if i >= args.warmup or nsamples == 1:
                vals.append(cb["value"])
            if time.time() - t0 > 200.0:
                break
        v = float(np.mean(vals)) if vals else cb["value"]
        cb["value"] = v
        cb["samples_timed"] = len(vals)
        out = dict(base, impl="reference", value=v, ms_per_step=1e3 / v, extrapolated=True,
                   config={"workload": desc, "elements": P.nelem, "dof": P.free_dofs(), "parallelism": "host CPU",
                           "cg_iters_per_step": hint, "cg_iters_source": src, "strip": cb["strip"],
                           "note": "value = 1 / (modelled seconds per full-size design iteration): element, assembly, optimiser and CG-iteration costs "
                                   "are MEASURED on the strip and scaled per element / per nonzero; ms_per_step is that modelled time, not wall time"},
                   cpu_baseline=cb, e2e={"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                   gpu_launches=0, wall_s=time.time() - t0)
        try:
            out["measured_pair"] = cpu_reference_real_iteration("pair" if args.workload != "c1" else "c1")
        except Exception as e:  # noqa: BLE001
            out["measured_pair"] = {"error": repr(e)[:200]}
        out["wall_s"] = time.time() - t0
        print(json.dumps(out))
        return 0

    import torch
    import torch.distributed as dist
    from pansfem2_b200 import capi

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    ctx = capi.Context(local_rank)
    env = {"ctx": ctx, "world": world, "rank": rank, "D": capi.Dist(ctx, rank, world) if world > 1 else None}
    K, W = args.steps, args.warmup
    t_start = time.time()
    record = {}

    # ---- headline workload: value, e2e, extra legs -----------------------------------------------------------------------
    R = Runner(env, args.workload)
    sampler = ClockSampler(local_rank)
    sampler.start()
    Lv = R.leg(K, W, tag="value")
    roof = R.roofline(Lv)
    Le = R.leg(K, W, host=True, tag="e2e")
    clocks = sampler.stop()
    launches = int(R.sum_over_ranks(Lv["launches"]))
    legs = {}
    if not args.no_extra_legs:
        Kx = min(K, 8)
        variants = [("warm_start", dict(warm_start=True)), ("matrix_free", dict(matrix_free=True)),
                    ("warm_start+matrix_free", dict(warm_start=True, matrix_free=True))]
        if world > 1 and os.environ.get("PF2_P2P", "1") != "0":
            variants.append(("single_reduction_cg", dict(single_reduction=True)))
            variants.append(("warm_start+matrix_free+single_reduction_cg", dict(warm_start=True, matrix_free=True, single_reduction=True)))
        for tag, kw in variants:
            try:
                Lx = R.leg(Kx, W, tag=tag, **kw)
                n = min(len(Lx["objective"]), len(Lv["objective"]))
                legs[tag] = brief(Lx, {"single_reduction_solves": Lx["pcg"].get("single_reduction_solves", 0),
                                       "objective_max_rel_diff_vs_value_leg": max(abs(a - b) / abs(b) for a, b in zip(Lx["objective"][:n], Lv["objective"][:n])),
                                       "value_leg_same_window": (Kx / (1e-3 * sum(s["phase_ms"]["solve"] + s["phase_ms"]["assemble"] + s["phase_ms"]["update"] +
                                                                                    s["phase_ms"]["filter"] + s["phase_ms"]["sens"] + s["phase_ms"]["filter_sens"]
                                                                                    for s in Lv["steps"][:Kx]))) if Kx <= len(Lv["steps"]) else None})
            except capi.Pf2Error as e:
                legs[tag] = {"unavailable": str(e)[:160]}
    nelem, P = R.nelem, R.P
    asm_ms = float(np.mean([st["phase_ms"]["assemble"] for st in Lv["steps"]]))
    out = dict(base, value=Lv["value"], ms_per_step=Lv["ms_per_step"],
               config={"workload": desc, "elements": R.nelem_global, "dof_local": R.S.A.rows, "nnz_local": R.S.A.nnz,
                       "parallelism": "1 GPU" if world == 1 else (f"{world} x-slabs (row-block partition; PCG halo + allreduce inside the persistent kernel over NVLink peer memory, "
                                                                    "NCCL for the per-design-iteration exchanges)" if os.environ.get("PF2_P2P", "1") != "0"
                                                                    else f"{world} x-slabs (row-block partition, NCCL halo exchange + allreduce)"),
                       "l2": "working set (CSR values+indices %.0f MB per rank) vs 126 MB L2; no flush between iterations (every iteration streams it)" % (12 * R.S.A.nnz / 1e6),
                       "cg_iters_per_step": Lv["cg_iters_per_step"], "solver": "ScalingCG eps=1e-10 x0=0", "operator": "assembled CSR (SELL-32 mirror)",
                       "design_iterations_timed": f"k = {W} .. {W + K - 1} after pf2_simp_reset (same window for value, e2e and the legs)"},
               e2e={"value": Le["value"], "unit": UNIT, "h2d_bytes_per_step": int(8 * nelem), "d2h_bytes_per_step": int(16 * nelem),
                    "ms_per_step": Le["ms_per_step"], "cg_iters_per_step": Le["cg_iters_per_step"]},
               gpu_launches=launches, clocks=clocks, roofline=roof, wall_s=Lv["wall"],
               phases_ms=Lv["steps"][-1]["phase_ms"], objective=Lv["objective"][W:], cg_relres_max=Lv["cg_relres_max"])
    if asm_ms > 0:
        bytes_per_elem = {2: 590.0, 1: 175.0, 3: 4300.0}.get(P.ndof, 0.0)     # Q4 plane strain / heat / hex8, map included
        out["assembly"] = {"elements_per_s": nelem / (asm_ms * 1e-3), "ms": asm_ms, "elements_local": int(nelem),
                           "algorithmic_bytes_per_element": bytes_per_elem, "algorithmic_GBps": bytes_per_elem * nelem / (asm_ms * 1e-3) / 1e9}
    if legs:
        out["legs"] = legs
    pv = R.parity_vs_n1()
    if pv:
        out["parity_vs_n1"] = pv
    record[args.workload] = {"objective": Lv["objective"], "cg_iters_by_k": Lv["cg_iters_by_k"], "warmup": W, "steps": K}
    cb_problem, cb_iters = P, Lv["cg_iters_per_step"]
    R.close()

    # ---- BASELINE.json's metric size: 2 M dof ----------------------------------------------------------------------------
    if not args.no_headline_2m and args.workload != "2m":
        try:
            R2 = Runner(env, "2m")
            K2 = min(K, 10)
            L2 = R2.leg(K2, W, tag="value")
            roof2 = R2.roofline(L2)
            L2e = R2.leg(K2, W, host=True, tag="e2e")
            out["headline_2m"] = dict(brief(L2), workload=WORKLOADS["2m"][4], warmup=W,
                                      e2e={"value": L2e["value"], "unit": UNIT, "ms_per_step": L2e["ms_per_step"], "h2d_bytes_per_step": int(8 * R2.nelem),
                                           "d2h_bytes_per_step": int(16 * R2.nelem), "cg_iters_per_step": L2e["cg_iters_per_step"]},
                                      roofline=roof2, phases_ms=L2["steps"][-1]["phase_ms"], objective=L2["objective"][W:], parity_vs_n1=R2.parity_vs_n1(),
                                      target=">= 1 design iteration/s on one B200 (BASELINE.json north_star)")
            record["2m"] = {"objective": L2["objective"], "cg_iters_by_k": L2["cg_iters_by_k"], "warmup": W, "steps": K2}
            if world == 1:
                try:
                    out["headline_2m"]["ilu0cg"] = R2.ilu_probe(ilu_itrmax=40)
                except Exception as e:  # noqa: BLE001
                    out["headline_2m"]["ilu0cg"] = {"error": repr(e)[:200]}
            R2.close()
        except Exception as e:  # noqa: BLE001
            out["headline_2m"] = {"error": repr(e)[:300]}

    # ---- hex8 configs, row-partitioned at N > 1 ---------------------------------------------------------------------------
    if not args.no_hex8 and args.workload not in ("c4", "c5"):
        hx = {}
        for name in [n for n in args.hex8.split(",") if n]:
            try:
                t0 = time.time()
                Rh = Runner(env, name)
                setup_s = time.time() - t0
                if name == "c5":
                    Lh = Rh.leg(2, 1, tag="value")             # design loop: k = 1, 2 timed after one warm-up iteration
                    what = "design iterations k = 1, 2 (k = 0 untimed)"
                else:
                    Lh = Rh.leg(2, 1, repeat_first=True, tag="value")      # 'single SIMP iteration': k = 0 from the uniform design, three times
                    what = "design iteration k = 0 from the uniform design, run 3 times (first untimed)"
                rh = Rh.roofline(Lh)
                hx[name] = dict(brief(Lh), workload=WORKLOADS[name][4], timed=what, dof_local=Rh.S.A.rows, nnz_local=Rh.S.A.nnz, setup_s=setup_s,
                                phases_ms=Lh["steps"][-1]["phase_ms"], objective=Lh["objective"], parity_vs_n1=Rh.parity_vs_n1(),
                                pcg_ms_per_cg_iteration=(rh or {}).get("ms_per_cg_iteration"), roofline_frac=(rh or {}).get("frac"),
                                spmv=(rh or {}).get("spmv"))
                record[name] = {"objective": Lh["objective"], "cg_iters_by_k": Lh["cg_iters_by_k"], "warmup": 1, "steps": 2}
                Rh.close()
            except Exception as e:  # noqa: BLE001
                hx[name] = {"error": repr(e)[:300]}
        out["hex8_scaling"] = hx

    # ---- one measured GPU / CPU pair at a size the reference finishes in seconds -------------------------------------------
    if world == 1 and not args.no_cpu_baseline:
        try:
            Rp = Runner(env, "pair")
            Lp = Rp.leg(1, 1, repeat_first=True, tag="value")
            cpu_pair = cpu_reference_real_iteration("pair")
            out["measured_pair"] = {"workload": WORKLOADS["pair"][4], "what": "design iteration k = 0, nothing modelled on either side",
                                    "gpu_seconds": Lp["ms"] * 1e-3, "gpu_cg_iters": Lp["cg_iters_per_step"], "gpu_objective": Lp["objective"][-1],
                                    "cpu": cpu_pair, "ratio": cpu_pair["seconds"] / (Lp["ms"] * 1e-3),
                                    "objective_rel_diff": abs(Lp["objective"][-1] - cpu_pair["objective"]) / abs(cpu_pair["objective"])}
            try:
                out["measured_pair"]["ilu0cg"] = Rp.ilu_probe()
            except Exception as e:  # noqa: BLE001
                out["measured_pair"]["ilu0cg"] = {"error": repr(e)[:200]}
            Rp.close()
        except Exception as e:  # noqa: BLE001
            out["measured_pair"] = {"error": repr(e)[:300]}
    out["total_wall_s"] = time.time() - t_start
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            try:
                out["cpu_baseline"] = cpu_reference_sample(cb_problem, cb_iters, "measured by this run's value leg")
            except Exception as e:  # the checker must never take the bench down
                out["cpu_baseline"] = {"error": repr(e)[:200]}
        if args.record:
            json.dump(record, open(args.record, "w"))
        print(json.dumps(out, default=float))
    if world > 1:
        dist.barrier()
        env["D"].close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
