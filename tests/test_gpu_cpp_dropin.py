"""The C++ side of the drop-in boundary on the GPU.

  * sample_optimize_density_batched : our driver on the batched, device-resident API (pansfem2_b200/src/B200/Batched.h)
  * dropin_density_oc / dropin_solid_linear : the reference's UNMODIFIED drivers (sample/optimize/sample_optimize_density_oc.cpp,
    sample/solid/sample_linear.cpp) compiled against the header mirror pansfem2_b200/src instead of the reference's src/
    (built by pansfem2_b200/build_cpp.py where /root/reference exists; the binaries travel to the GPU box).
Outputs are the reference's own VTK files, compared with the committed goldens at their 6 printed digits.
"""
import os
import subprocess

import numpy as np
import pytest


pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "pansfem2_b200", "bin")


def parse_vtk(path):
    lines = open(path).read().split("\n")
    out, i = {}, 0
    while i < len(lines):
        ln = lines[i]
        if ln.startswith("VECTORS") or ln.startswith("SCALARS"):
            name = ln.split()[1]
            i += 1 if ln.startswith("VECTORS") else 2
            rows = []
            while i < len(lines) and lines[i].strip() and not lines[i][0].isalpha():
                rows.append([float(v) for v in lines[i].split()])
                i += 1
            out[name] = np.array(rows).squeeze()
            continue
        i += 1
    return out


def need(name):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (python -m pansfem2_b200.build_cpp)")
    return exe


@pytest.mark.parametrize("opt,tag,iters", [("oc", "oc", 66), ("mma", "mma", 56), ("conlin", "conlin", 133)])
def test_batched_driver_reproduces_golden_vtk(tmp_path, golden_dir, opt, tag, iters):
    exe = need("sample_optimize_density_batched")
    out = tmp_path / "result.vtk"
    r = subprocess.run([exe, opt, "60", "40", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Optimized" in r.stdout and f"k = {iters - 1}\t" in r.stdout and f"k = {iters}\t" not in r.stdout
    got, g = parse_vtk(out), np.load(os.path.join(golden_dir, f"density_{tag}.npz"))
    assert np.abs(got["s"] - g["rho"]).max() < 2e-6
    np.testing.assert_allclose(got["u"][:, :2], g["u"], rtol=1e-5, atol=1e-11)
    np.testing.assert_allclose(got["r"][:, :2], g["r"], rtol=1e-5, atol=1e-7)
    # pf2_simp_export_vtk (fields staged on the device, C writer) against the mirror's ExportToVTK.h writers (pinned byte for byte to the
    # reference's in tests/test_io_formats.py) on the same state
    assert open(str(out) + ".device.vtk", "rb").read() == open(str(out) + ".nor.vtk", "rb").read()
    dev = parse_vtk(str(out) + ".device_r.vtk")               # with the reactions recomputed by the export: same fields as the golden file
    np.testing.assert_allclose(dev["u"][:, :2], g["u"], rtol=1e-5, atol=1e-11)
    np.testing.assert_allclose(dev["r"][:, :2], g["r"], rtol=1e-5, atol=1e-7)
    assert np.abs(dev["s"] - g["rho"]).max() < 2e-6


def test_unmodified_reference_oc_driver_on_the_header_mirror(tmp_path, golden_dir):
    exe = need("dropin_density_oc")
    (tmp_path / "sample" / "optimize").mkdir(parents=True)
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Optimized" in r.stdout
    files = sorted((tmp_path / "sample" / "optimize").glob("result*.vtk"), key=lambda p: int(p.stem[6:]))
    assert len(files) == 66                       # k = 0..65, as the reference run
    got, g = parse_vtk(files[-1]), np.load(os.path.join(golden_dir, "density_oc.npz"))
    assert np.abs(got["s"] - g["rho"]).max() < 2e-6
    np.testing.assert_allclose(got["u"][:, :2], g["u"], rtol=1e-5, atol=1e-11)
    np.testing.assert_allclose(got["r"][:, :2], g["r"], rtol=1e-5, atol=1e-7)


def test_unmodified_reference_conlin_driver_on_the_header_mirror(tmp_path, golden_dir):
    """sample_optimize_density_CONLIN.cpp, unmodified, on CONLIN.h / HeavisideFilter.h of the mirror: 133 iterations."""
    exe = need("dropin_density_conlin")
    (tmp_path / "sample" / "optimize").mkdir(parents=True)
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-2000:]
    files = sorted((tmp_path / "sample" / "optimize").glob("result*.vtk"), key=lambda p: int(p.stem[6:]))
    assert len(files) == 133                      # k = 0..132, as the reference run
    got, g = parse_vtk(files[-1]), np.load(os.path.join(golden_dir, "density_conlin.npz"))
    assert np.abs(got["s"] - g["rho"]).max() < 2e-6
    np.testing.assert_allclose(got["u"][:, :2], g["u"], rtol=1e-5, atol=1e-11)
    np.testing.assert_allclose(got["r"][:, :2], g["r"], rtol=1e-5, atol=1e-7)


def test_unmodified_reference_solid_driver_on_the_header_mirror(tmp_path, golden_dir):
    exe = need("dropin_solid_linear")
    g = np.load(os.path.join(golden_dir, "solid_linear.npz"))
    d = tmp_path / "sample" / "solid"
    d.mkdir(parents=True)
    with open(d / "Node.csv", "w") as f:
        f.write("ID,x0,x1,x2\n")
        for i, c in enumerate(g["coords"]):
            f.write(f"{i},{c[0]:.17g},{c[1]:.17g},{c[2]:.17g}\n")
    with open(d / "Element.csv", "w") as f:
        f.write("ID,n0,n1,n3,n4,n5,n6,n7,n8\n")
        for i, e in enumerate(g["conn"]):
            f.write(f"{i}," + ",".join(str(int(v)) for v in e) + "\n")
    for name, (nn, dd, vv), hdr in (("Dirichlet.csv", (g["fix_node"], g["fix_dof"], g["fix_val"]), "idn,u,v,w"),
                                    ("Neumann.csv", (g["load_node"], g["load_dof"], g["load_val"]), "idn,fx,fy,fz")):
        rows = {}
        for n, dof, v in zip(nn, dd, vv):
            rows.setdefault(int(n), ["free"] * 3)[int(dof)] = f"{v:.17g}"
        with open(d / name, "w") as f:
            f.write(hdr + "\n")
            for n in sorted(rows):
                f.write(f"{n}," + ",".join(rows[n]) + "\n")
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got = parse_vtk(d / "result_linear.vtk")
    assert abs(np.abs(got["u"]).max() - 1.48963) < 1e-5
    np.testing.assert_allclose(got["u"], g["u"], rtol=6e-6, atol=1e-9)


def test_unmodified_reference_planestrain_t3_driver_on_the_header_mirror(tmp_path, golden_dir):
    """sample/planestrain/sample_planestrain.cpp, unmodified: PlaneStrainStiffness<ShapeFunction3Triangle, Gauss1Triangle> on the
    device, body / surface force vectors through the mirror's host routines, CG; output against the committed result.vtk."""
    exe = need("dropin_planestrain_t3")
    (tmp_path / "sample" / "planestrain").mkdir(parents=True)
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got, g = parse_vtk(tmp_path / "sample" / "planestrain" / "result.vtk"), np.load(os.path.join(golden_dir, "t3_samples.npz"))
    np.testing.assert_allclose(got["u"][:, :2], g["ps_u"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(got["r"][:, :2], g["ps_r"], rtol=1e-5, atol=2e-3)


def test_batched_planestrain_driver_with_device_load_vectors(tmp_path, golden_dir):
    """sample_planestrain.cpp through the batched API: B200::AssembleBatched + B200::AssembleLoadVector (body force on all T3 elements,
    traction on the loaded edges, both integrated and assembled on the device with the sample's own functors) -> the committed result.vtk."""
    exe = need("sample_planestrain_batched")
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    u = np.array([[float(v) for v in ln.split()[2:4]] for ln in r.stdout.splitlines() if ln.startswith("u ")])
    g = np.load(os.path.join(golden_dir, "t3_samples.npz"))
    np.testing.assert_allclose(u, g["ps_u"], rtol=1e-5, atol=1e-9)


def test_unmodified_reference_homogenization_driver_on_the_header_mirror(tmp_path, golden_dir):
    """sample/homogenization/sample_homogenization.cpp, unmodified: SquareAnnulusMesh2, ImportPeriodicFromCSV + SetPeriodic, per-element
    PlaneStrainStiffness on the device, HomogenizePlaneStrainBodyForce / WeakSpring / ...Constitutive on the host, three ScalingCG solves on
    the device -> the committed result_microscopic.vtk and the matrices the reference run prints.  (Added after the last GPU slot of r01;
    tests/test_homogenization_pinned.py is the CPU replay that set the tolerances.)"""
    exe = need("dropin_homogenization")
    g = np.load(os.path.join(golden_dir, "homogenization.npz"))
    d = tmp_path / "sample" / "homogenization"
    d.mkdir(parents=True)
    with open(d / "Periodic.csv", "w") as f:
        f.write("masterid,slaveid\n")
        for m, s_ in g["pairs"]:
            f.write(f"{int(m)},{int(s_)}\n")
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got = parse_vtk(d / "result_microscopic.vtk")
    for c in range(3):
        assert np.abs(got[f"chi{c}"][:, :2] - g[f"chi{c}"]).max() < 2e-6
    vals = np.array([float(v) for v in r.stdout.split() if v[0] in "-0123456789."])[-18:].reshape(2, 3, 3)
    np.testing.assert_allclose(vals[0], g["check"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(vals[1], g["CH"], rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("family,nx,ny", [("t3", 12, 8), ("q8sri", 10, 6)])
def test_batched_driver_on_other_element_families(family, nx, ny):
    """sample_optimize_density_families: the batched C++ API with PlaneStressStiffnessTag<3Triangle, Gauss1Triangle> and
    PlaneStrainStiffnessSRITag<8Square, Gauss4Square, Gauss9Square>; the run is replayed in the oracle on the same mesh."""
    from oracle import portlib as orc
    from pansfem2_b200 import eqcode as ec, mesher, problems
    exe = need("sample_optimize_density_families")
    r = subprocess.run([exe, family, str(nx), str(ny), "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    hist = np.array([[float(t) for t in (ln.split("\t")[2], ln.split("\t")[4])] for ln in r.stdout.strip().split("\n") if ln.startswith("k =")])
    assert hist.shape == (4, 2)
    coords, quads = mesher.square_mesh(float(nx), float(ny), nx, ny)
    if family == "t3":
        conn = np.stack([quads[:, [1, 2, 0]], quads[:, [2, 3, 0]]], axis=1).reshape(-1, 3)
        eq = ec.eq_code(ec.PHYS_PLANESTRESS, ec.SHAPE_T3, ec.QUAD_G1TRI)
    else:
        mid, extra, conn = {}, [], []
        for q in quads:
            e = list(q)
            for a in range(4):
                key = (min(q[a], q[(a + 1) % 4]), max(q[a], q[(a + 1) % 4]))
                if key not in mid:
                    mid[key] = len(coords) + len(extra)
                    extra.append((coords[key[0]] + coords[key[1]]) / 2.0)
                e.append(mid[key])
            conn.append(e)
        coords, conn = np.vstack([coords, np.array(extra)]), np.array(conn, np.int32)
        eq = ec.eq_code(ec.PHYS_PLANESTRAIN_SRI, ec.SHAPE_Q8, ec.QUAD_G9SQ, ec.QUAD_G4SQ)
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    fixed = mesher.fixed_list(coords, [0, 1], lambda x: np.abs(x[:, 0]) < 1e-9)
    loads = mesher.fixed_list(coords, [1], lambda x: (np.abs(x[:, 0] - nx) < 1e-9) & (np.abs(x[:, 1] - ny / 2) < 1e-9), -1.0)
    nbrs = mesher.filter_neighbors_allpairs(mesher.element_centroids(coords, conn), 1.5)
    P = problems.Problem("replay", eq, coords, conn, fixed, loads, nbrs, (nx, ny), filter_kind=problems.FILTER_DENSITY, beta_period=0)
    R = orc.simp_run(eq, P.coords, P.conn, P.fixed, P.loads, P.filter_kind, P.nbrs, P.opt_kind, P.optp(), P.params(), 4,
                     np.full(P.nelem, 0.5), check_convergence=False)
    np.testing.assert_allclose(hist[:, 0], R["hist"][:, 0], rtol=1e-8)
    np.testing.assert_allclose(hist[:, 1], R["hist"][:, 1], rtol=0, atol=1e-9)


def _levelset_stdout(txt):
    import re
    return np.array([[float(v) for v in re.findall(r"= ([-+0-9.e]+)", ln)] for ln in txt.split("\n") if ln.startswith("t =")])


def test_batched_levelset_driver_reproduces_the_reference_run(tmp_path, golden_dir):
    """sample_optimize_levelset_batched (B200::LevelSetLoop): the console history and the final fields of the reference's
    sample_optimize_levelset.cpp run (tests/golden/levelset.npz: stdout of the unmodified sample + its last VTK)."""
    exe = need("sample_optimize_levelset_batched")
    out = tmp_path / "result.vtk"
    r = subprocess.run([exe, "60", "40", str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    g = np.load(os.path.join(golden_dir, "levelset.npz"))
    hist = _levelset_stdout(r.stdout)
    assert "Convergence" in r.stdout and hist.shape == g["stdout_hist"].shape == (115, 4)
    np.testing.assert_allclose(hist, g["stdout_hist"], rtol=2e-5)             # both sides print 6 digits
    got = parse_vtk(out)
    assert np.array_equal(got["str"], g["vtk_str"])
    np.testing.assert_allclose(got["phi"], g["vtk_phi"], rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(got["u"][:, :2], g["vtk_u"], rtol=2e-5, atol=1e-9)


def test_unmodified_reference_levelset_driver_on_the_header_mirror(tmp_path, golden_dir):
    """sample/optimize/sample_optimize_levelset.cpp, unmodified, on the mirror's PlaneStress.h / ReactionDiffusion.h / General.h
    (per-element device calls + host containers): same history, same convergence iteration, same final structure."""
    exe = need("dropin_levelset")
    (tmp_path / "sample" / "optimize").mkdir(parents=True)
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-2000:]
    g = np.load(os.path.join(golden_dir, "levelset.npz"))
    hist = _levelset_stdout(r.stdout)
    assert "Convergence" in r.stdout and hist.shape == (115, 4)
    np.testing.assert_allclose(hist, g["stdout_hist"], rtol=2e-5)
    files = sorted((tmp_path / "sample" / "optimize").glob("result*.vtk"), key=lambda p: int(p.stem[6:]))
    assert len(files) == int(g["vtk_count"]) == 116
    got = parse_vtk(files[-1])
    assert np.array_equal(got["str"], g["vtk_str"])
    np.testing.assert_allclose(got["phi"], g["vtk_phi"], rtol=2e-5, atol=1e-9)
