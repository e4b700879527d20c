"""One-screen digest of a bench.py JSON line: python tools/bench_brief.py file.json"""
import json
import sys

txt = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")]
if not txt:
    print("no JSON line in", sys.argv[1]); sys.exit(0)
d = json.loads(txt[-1])
cg = d["config"]["cg_iters_per_step"]
r = d.get("roofline") or {}
print(f"N={d['n_gpus']} value={d['value']:.4f} e2e={d['e2e']['value']:.4f} cg/step={cg:.0f} ms/cg_iter={d['ms_per_step'] / cg:.5f} roof_frac={r.get('frac')} "
      f"kernel={str(r.get('kernel'))[:28]} parity={d.get('parity_vs_n1')}")
sp = r.get("spmv") or r.get("pcg_iteration")
print("   spmv/phases:", {k: (round(v, 5) if isinstance(v, float) else v) for k, v in (sp or {}).items() if k not in ("note", "alone_note")})
for k, v in (d.get("legs") or {}).items():
    print("   leg", k, {a: (round(b, 5) if isinstance(b, float) else b) for a, b in v.items()})
h = d.get("headline_2m")
if h:
    print("   2m:", {k: (round(v, 5) if isinstance(v, float) else v) for k, v in h.items() if k in ("value", "ms_per_step", "cg_iters_per_step", "error")}, "e2e", (h.get("e2e") or {}).get("value"))
for k, v in (d.get("hex8_scaling") or {}).items():
    print("   hex8", k, {a: (round(b, 5) if isinstance(b, float) else b) for a, b in v.items() if a in ("value", "ms_per_step", "cg_iters_per_step", "pcg_ms_per_cg_iteration", "roofline_frac", "setup_s", "error", "parity_vs_n1")})
mp = d.get("measured_pair")
if mp:
    print("   pair:", {k: mp.get(k) for k in ("gpu_seconds", "ratio", "objective_rel_diff", "error")})
    print("   pair ilu:", mp.get("ilu0cg"))
if h and h.get("ilu0cg"):
    print("   2m ilu:", h.get("ilu0cg"))
cb = d.get("cpu_baseline")
if cb:
    print("   cpu_baseline:", {k: cb.get(k) for k in ("value", "cores", "extrapolated", "strip", "error")})
print("   total_wall_s", d.get("total_wall_s"), "clocks", d.get("clocks"))
