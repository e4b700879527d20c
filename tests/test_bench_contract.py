"""bench.py's reference arm (`--impl reference`) runs on the host cores only, so its JSON contract can be checked without a GPU:
one line, the metric / unit / config of the native arm, `impl`, `cpu_baseline` and a zero-copy `e2e` block."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().split("\n") if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "SIMP design iterations/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and "workload" in d["config"] and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    # the arm scales a strip to the full mesh and says so; one design iteration is run for real next to it
    assert d["extrapolated"] is True and cb["extrapolated"] is True and cb["strip"] and "EXTRAPOLATED" in cb["sample"]
    assert d["config"]["cg_iters_source"] and d["measured_pair"]["extrapolated"] is False and d["measured_pair"]["seconds"] > 0


def test_reference_arm_uses_the_cg_count_the_native_arm_measured():
    sys.path.insert(0, ROOT)
    import bench
    P = bench.make_problem("c1")
    n, src = bench.reference_cg_count("c2", 5, 20, bench.make_problem("pair"))
    rec = json.load(open(os.path.join(ROOT, "profiles", "measured_cg_iters.json")))["c2"]["cg_iters_by_k"]
    assert abs(n - sum(rec[5:25]) / 20.0) < 1e-9 and "measured by the native arm" in src
    n2, src2 = bench.reference_cg_count("c3", 5, 20, P)          # no record for that workload: the growth law, labelled as an estimate
    assert "ESTIMATED" in src2 and n2 > 0


def test_committed_n1_objective_histories_cover_the_bench_legs():
    obj = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_objective_n1.json")))
    assert len(obj["c2"]) >= 25 and len(obj["2m"]) >= 15 and len(obj["c4"]) >= 1 and len(obj["c5"]) >= 3
    assert all(b < a for a, b in zip(obj["c2"][:10], obj["c2"][1:11]))          # the compliance falls monotonically over the first iterations
