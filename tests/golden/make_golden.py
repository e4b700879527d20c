"""Generate tests/golden/*.npz from the reference tree (run in the build container where /root/reference exists):

    python tests/golden/make_golden.py

Sources (all under /root/reference, nothing is copied verbatim - values are parsed into arrays):
  * sample/optimize/Density_OC.vtk  / Density_MMA.vtk / Density_CONLIN.vtk : committed final-iteration outputs of
    sample_optimize_density_{oc,mma,CONLIN}.cpp (VECTORS u, VECTORS r, SCALARS s = filtered rho), 6 significant digits.
  * sample/solid/result_linear.vtk + {Node,Element,Dirichlet,Neumann}.csv : hex8 50x5x5 cantilever solved by
    sample/solid/sample_linear.cpp (SolidLinearIsotropicElastic + ScalingCG).
  * src/Optimize/Solver/test_MMA.cpp, test_MMA_3.cpp, test_MMA_TOY2.cpp : the print-only known-answer mains are
    compiled UNMODIFIED and their stdout parsed (iterates per outer iteration).
  * live reference (oracle/_ref/libpf2ref.so): unit-element matrices, C1 iteration history, a 12x8 system's CSR,
    ILU0 factors and solver outputs - the fixtures the GPU box checks the C restatement against.

Sub-commands (python tests/golden/make_golden.py <name>) regenerate one family of fixtures:
    conlin | families | levelset | krylov     the widened rows 1-4 (live_conlin, live_families + t3_samples + shape_tables, levelset, live_krylov)
    advection                                  Advection.h family: 180 element cases, both advection samples, sample/heattransfer/dynamic.vtk
    plane_d | homogenization                   PlaneStiffness* with a caller-supplied D; sample/homogenization (pairs, result_microscopic.vtk, stdout)
    io | meshers | routines | linalg           host-side mirror: small C++ programs under tests/cpp built against the REFERENCE's headers, stdout kept
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import reflib  # noqa: E402
from pansfem2_b200 import problems  # noqa: E402


def parse_vtk_fields(path):
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


def csv_rows(path):
    rows = [r.strip().split(",") for r in open(path).read().strip().split("\n")[1:]]
    return rows


def golden_simp():
    for tag in ("OC", "MMA", "CONLIN"):
        f = parse_vtk_fields(f"{REF}/sample/optimize/Density_{tag}.vtk")
        np.savez_compressed(f"{OUT}/density_{tag.lower()}.npz", u=f["u"][:, :2], r=f["r"][:, :2], rho=f["s"])
        print(tag, {k: v.shape for k, v in f.items()})


def golden_solid():
    nodes = np.array([[float(v) for v in r[1:4]] for r in csv_rows(f"{REF}/sample/solid/Node.csv")])
    elems = np.array([[int(v) for v in r[1:9]] for r in csv_rows(f"{REF}/sample/solid/Element.csv")], dtype=np.int32)
    fn, fd, fv = [], [], []
    for r in csv_rows(f"{REF}/sample/solid/Dirichlet.csv"):
        for d, tok in enumerate(r[1:4]):
            if tok != "free":
                fn.append(int(r[0])); fd.append(d); fv.append(float(tok))
    ln, ld, lv = [], [], []
    for r in csv_rows(f"{REF}/sample/solid/Neumann.csv"):
        for d, tok in enumerate(r[1:4]):
            if tok != "free":
                ln.append(int(r[0])); ld.append(d); lv.append(float(tok))
    f = parse_vtk_fields(f"{REF}/sample/solid/result_linear.vtk")
    np.savez_compressed(f"{OUT}/solid_linear.npz", coords=nodes, conn=elems,
                        fix_node=np.array(fn, np.int32), fix_dof=np.array(fd, np.int32), fix_val=np.array(fv),
                        load_node=np.array(ln, np.int32), load_dof=np.array(ld, np.int32), load_val=np.array(lv),
                        u=f["u"])
    print("solid", nodes.shape, elems.shape, len(fn), len(ln), np.abs(f["u"]).max())


def golden_mma_kat():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name in ("test_MMA", "test_MMA_3", "test_MMA_TOY2"):
            exe = os.path.join(tmp, name)
            subprocess.run(["g++", "-O3", "-fopenmp", f"{REF}/src/Optimize/Solver/{name}.cpp", "-o", exe], check=True)
            txt = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
            rows = []
            for ln in txt.split("\n"):
                if ln.startswith("k ="):
                    nums = [float(v) for v in re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", ln)]
                    rows.append(nums)
            width = max(len(r) for r in rows)
            out[name] = np.array([r for r in rows if len(r) == width])
            print(name, out[name].shape, out[name][-1])
    np.savez_compressed(f"{OUT}/mma_kat.npz", **out)


def golden_live():
    """Fixtures produced by the live reference so the GPU box (no /root/reference) can still pin the C restatement."""
    reflib.set_num_threads(1)
    d = {}
    # unit elements and a distorted one each
    q4 = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], float)
    q4d = np.array([[0.1, -0.2], [1.3, 0.1], [1.1, 1.4], [-0.2, 0.9]], float)
    h8 = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    rng = np.random.default_rng(20201017)
    h8d = h8 * np.array([1.3, 0.8, 1.1]) + 0.15 * rng.uniform(-1, 1, h8.shape)
    d["q4"], d["q4d"], d["h8"], d["h8d"] = q4, q4d, h8, h8d
    d["ke_ps_q4"] = reflib.element_matrix(reflib.EQ_PLANESTRAIN, q4, 1.0, 0.3, 1.0)
    d["ke_ps_q4d"] = reflib.element_matrix(reflib.EQ_PLANESTRAIN, q4d, 2.5, 0.3, 0.7)
    d["ke_heat_q4"] = reflib.element_matrix(reflib.EQ_HEAT, q4, 1.0, 0.0, 1.0)
    d["ke_heat_q4d"] = reflib.element_matrix(reflib.EQ_HEAT, q4d, 2.5, 0.0, 0.7)
    d["ke_solid_h8"] = reflib.element_matrix(reflib.EQ_SOLID, h8, 1.0, 0.3, 1.0)
    d["ke_solid_h8d"] = reflib.element_matrix(reflib.EQ_SOLID, h8d, 2.5, 0.3, 1.0)
    # small plane-strain system with a non-zero Dirichlet value (exercises the F lift, Assembling.h:59)
    P = problems.cantilever2d(12, 8)
    fixed = (P.fixed[0], P.fixed[1], np.where(P.fixed[1] == 0, 0.01, -0.02))
    Emod = rng.uniform(0.5, 2.0, P.nelem)
    S = reflib.assemble(P.eq, P.coords, P.conn, fixed, P.loads, Emod)
    indptr, indices, data, F = S.arrays()
    d.update(sys_Emod=Emod, sys_fixval=fixed[2], sys_indptr=indptr, sys_indices=indices, sys_data=data, sys_F=F)
    for kind, nm in ((0, "cg"), (1, "scalingcg"), (2, "ilu0cg")):
        d[f"sys_x_{nm}"] = S.solve(kind, F)[0]
    M = S.ilu0()
    d["sys_ilu0_data"] = M.arrays()[2]
    d["sys_preilu0"] = M.preilu0(F)
    # filters / OC on the same mesh
    s = rng.uniform(0.05, 1.0, P.nelem)
    dfdrho = -rng.uniform(0.1, 2.0, P.nelem)
    for kind, nm in ((reflib.FILTER_DENSITY, "density"), (reflib.FILTER_HEAVISIDE, "heaviside")):
        flt = reflib.Filter(kind, *P.nbrs)
        d[f"flt_{nm}_rho"] = flt.apply(2.0, s)
        d[f"flt_{nm}_sens"] = flt.sens(2.0, s, dfdrho)
        oc = reflib.OC(P.nelem, *P.oc)
        dgds = flt.sens(2.0, s, np.full(P.nelem, 1.0 / (0.5 * P.nelem)))
        d[f"oc_{nm}_x"] = oc.update(flt, 2.0, 0.5, 1.0, s, 1.0, d[f"flt_{nm}_sens"], 0.1, dgds)
        d[f"oc_{nm}_dgds"] = dgds
    d["flt_s"], d["flt_dfdrho"] = s, dfdrho
    # C1 history (objective/constraint per design iteration) for OC and MMA, first 12 iterations + final state
    for opt, nm in ((problems.OPT_OC, "oc"), (problems.OPT_MMA, "mma")):
        P1 = problems.cantilever2d(60, 40, opt_kind=opt)
        R = reflib.simp_run(P1.eq, P1.coords, P1.conn, P1.fixed, P1.loads, P1.filter_kind, P1.nbrs, P1.opt_kind,
                            P1.optp(), P1.params(), 12, np.full(P1.nelem, 0.5), check_convergence=False)
        d[f"c1_{nm}_hist"] = R["hist"][:, :2]
        d[f"c1_{nm}_s12"] = R["s"]
        d[f"c1_{nm}_rho12"] = R["rho"]
        print("c1", nm, R["hist"][:3, :2], R["hist"][-1, :2])
    np.savez_compressed(f"{OUT}/live_reference.npz", **d)
    print("live fixtures:", len(d))


def golden_conlin():
    """CONLIN<T> and SensitivityFilter/SensitivityFilter2 fixtures from the live reference."""
    reflib.set_num_threads(1)
    d = {}
    rng = np.random.default_rng(20200719)
    P = problems.cantilever2d(12, 8)
    n = P.nelem
    s = rng.uniform(0.05, 1.0, n)
    dfds = -rng.uniform(0.1, 2.0, n)
    d["s"], d["dfds"] = s, dfds
    d["sens_sigmund"] = reflib.sensitivity_filter(2, P.nbrs, s, dfds)
    d["sens_borrvall"] = reflib.sensitivity_filter(3, P.nbrs, s, dfds)
    # three CONLIN updates with one and with two constraints (mixed-sign gradients exercise both p and q)
    for m in (1, 2):
        opt = reflib.CONLIN(n, m, 1.0, np.zeros(m), np.full(m, 1.0e4), np.zeros(m), 0.01, 1.0)
        opt.set_parameters(0.2, 1.0e-6)
        x = s.copy()
        xs = []
        for it in range(3):
            df = dfds * (1.0 + 0.1 * it) * np.where(np.arange(n) % 7 == 3, -0.05, 1.0)
            g = np.array([x.sum() / (0.5 * n) - 1.0, 0.3 - x[: n // 2].sum() / n][:m])
            dg = np.stack([np.full(n, 1.0 / (0.5 * n)), np.where(np.arange(n) < n // 2, -1.0 / n, 0.0)][:m])
            x = opt.update(x, 1.0, df, g, dg)
            xs.append(x.copy())
        d[f"conlin_m{m}_x"] = np.stack(xs)
    # C1 with the CONLIN driver: first 12 design iterations
    P1 = problems.cantilever2d(60, 40, opt_kind=problems.OPT_CONLIN)
    R = reflib.simp_run(P1.eq, P1.coords, P1.conn, P1.fixed, P1.loads, P1.filter_kind, P1.nbrs, P1.opt_kind,
                        P1.optp(), P1.params(), 12, np.full(P1.nelem, 0.5), check_convergence=False)
    d["c1_conlin_hist"] = R["hist"][:, :2]
    d["c1_conlin_s12"] = R["s"]
    d["c1_conlin_rho12"] = R["rho"]
    print("c1 conlin", R["hist"][:3, :2], R["hist"][-1, :2])
    np.savez_compressed(f"{OUT}/live_conlin.npz", **d)
    print("conlin fixtures:", len(d))


def parse_vtk_mesh(path):
    lines = open(path).read().split("\n")
    i = next(k for k, ln in enumerate(lines) if ln.startswith("POINTS"))
    npts = int(lines[i].split()[1])
    pts = np.array([[float(v) for v in lines[i + 1 + k].split()] for k in range(npts)])
    j = next(k for k, ln in enumerate(lines) if ln.startswith("CELLS"))
    ncell = int(lines[j].split()[1])
    cells = [[int(v) for v in lines[j + 1 + k].split()][1:] for k in range(ncell)]
    return pts, np.array(cells, dtype=np.int32)


def all_selections():
    from pansfem2_b200 import eqcode as ec
    out = []
    for phys in (ec.PHYS_PLANESTRAIN, ec.PHYS_PLANESTRESS, ec.PHYS_HEAT, ec.PHYS_PLANESTRAIN_SRI, ec.PHYS_SOLID, ec.PHYS_MASS,
                 ec.PHYS_PLANESTRAIN_BBAR, ec.PHYS_MASS2):
        shapes = (ec.SHAPE_TET4, ec.SHAPE_HEX8, ec.SHAPE_HEX20) if phys == ec.PHYS_SOLID else (ec.SHAPE_T3, ec.SHAPE_T6, ec.SHAPE_Q4, ec.SHAPE_Q8)
        for shape in shapes:
            for quad in ec.SHAPE_RULES[shape]:
                for q2 in (ec.SHAPE_RULES[shape] if phys in (ec.PHYS_PLANESTRAIN_SRI, ec.PHYS_PLANESTRAIN_BBAR) else (0,)):
                    out.append(ec.eq_code(phys, shape, quad, q2))
    # Wilson-Taylor incompatible modes: quadrilaterals, rules with off-centre points (the modes vanish at the centre)
    for shape in (ec.SHAPE_Q4, ec.SHAPE_Q8):
        for quad in (ec.QUAD_G4SQ, ec.QUAD_G9SQ):
            out.append(ec.eq_code(ec.PHYS_PLANESTRAIN_WT, shape, quad))
    return out


def golden_families():
    """Element families beyond Q4 / hex8: every <Equation, SF, IC> selection of the reference on a
    distorted element, assembled systems + solves + short SIMP runs on family meshes (live reference), and the reference's
    committed T3 outputs sample/heattransfer/static.vtk and sample/planestrain/result.vtk."""
    from pansfem2_b200 import eqcode as ec, mesher
    reflib.set_num_threads(1)
    rng = np.random.default_rng(20210301)
    d = {}
    sel = all_selections()
    d["selections"] = np.array(sel, np.int64)
    for eq in sel:
        shape = ec.fields(eq)[1]
        nat = mesher.NATURAL_NODES[ec.SHAPE_NAME[shape]]
        xe = nat * (np.array([1.3, 0.9]) if nat.shape[1] == 2 else np.array([1.3, 0.8, 1.1])) + 0.08 * rng.uniform(-1, 1, nat.shape)
        d[f"xe_{eq}"] = xe
        d[f"ke_{eq}"] = reflib.element_matrix(eq, xe, 2.5, 0.3, 0.7)
    # assembled systems (non-zero Dirichlet values) and ScalingCG solutions on small family meshes
    cases = {"t3_heat": (ec.eq_code(ec.PHYS_HEAT, ec.SHAPE_T3), (6, 4)),
             "t6_pstress": (ec.eq_code(ec.PHYS_PLANESTRESS, ec.SHAPE_T6, ec.QUAD_G3TRI), (5, 3)),
             "q8_sri": (ec.eq_code(ec.PHYS_PLANESTRAIN_SRI, ec.SHAPE_Q8, ec.QUAD_G9SQ, ec.QUAD_G4SQ), (5, 3)),
             "q8_pstrain": (ec.eq_code(ec.PHYS_PLANESTRAIN, ec.SHAPE_Q8, ec.QUAD_G9SQ), (4, 3)),
             "q4_wt": (ec.eq_code(ec.PHYS_PLANESTRAIN_WT, ec.SHAPE_Q4, ec.QUAD_G4SQ), (6, 4)),
             "q4_bbar": (ec.eq_code(ec.PHYS_PLANESTRAIN_BBAR, ec.SHAPE_Q4, ec.QUAD_G4SQ, ec.QUAD_G1SQ), (6, 4)),
             "tet4": (ec.eq_code(ec.PHYS_SOLID, ec.SHAPE_TET4), (3, 2, 2)),
             "hex20": (ec.eq_code(ec.PHYS_SOLID, ec.SHAPE_HEX20, ec.QUAD_G27CUBE), (3, 2, 2))}
    for nm, (eq, n) in cases.items():
        P = problems.family_problem(eq, n)
        fixed = (P.fixed[0], P.fixed[1], np.where(P.fixed[1] == 0, 0.01, -0.02))
        Emod = rng.uniform(0.5, 2.0, P.nelem)
        S = reflib.assemble(P.eq, P.coords, P.conn, fixed, P.loads, Emod, 0.3, 0.8)
        indptr, indices, data, F = S.arrays()
        d[f"{nm}_eq"], d[f"{nm}_n"] = np.int64(eq), np.array(n)
        d.update({f"{nm}_Emod": Emod, f"{nm}_indptr": indptr, f"{nm}_indices": indices, f"{nm}_data": data, f"{nm}_F": F,
                  f"{nm}_x": S.solve(1, F)[0]})
        # 4 SIMP iterations (OC + density filter), history and final design
        R = reflib.simp_run(P.eq, P.coords, P.conn, P.fixed, P.loads, P.filter_kind, P.nbrs, P.opt_kind, P.optp(), P.params(), 4,
                            np.full(P.nelem, 0.5), check_convergence=False)
        d[f"{nm}_hist"], d[f"{nm}_s4"] = R["hist"][:, :2], R["s"]
        print(nm, ec.describe(eq), P.nelem, "rows", S.rows, "f", R["hist"][:, 0])
    np.savez_compressed(f"{OUT}/live_families.npz", **d)
    # committed T3 outputs of the reference
    pts, cells = parse_vtk_mesh(f"{REF}/sample/heattransfer/static.vtk")
    f = parse_vtk_fields(f"{REF}/sample/heattransfer/static.vtk")
    T = f["T"] if f["T"].ndim == 1 else f["T"][:, 0]
    pts2, cells2 = parse_vtk_mesh(f"{REF}/sample/planestrain/result.vtk")
    f2 = parse_vtk_fields(f"{REF}/sample/planestrain/result.vtk")
    # N / dNdr / Points / Weights of every policy class, printed by the reference's own headers
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "shape_tables")
        subprocess.run(["g++", "-O1", "-std=c++17", f"-I{REF}/src", f"{ROOT}/tests/cpp/shape_tables.cpp", "-o", exe], check=True)
        open(f"{OUT}/shape_tables.txt", "w").write(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)
    np.savez_compressed(f"{OUT}/t3_samples.npz", heat_coords=pts[:, :2], heat_conn=cells, heat_T=T,
                        ps_coords=pts2[:, :2], ps_conn=cells2, ps_u=f2["u"][:, :2], ps_r=f2["r"][:, :2])
    print("t3 goldens", pts.shape, cells.shape, T.min(), T.max(), f2["u"])


def golden_levelset():
    """Level-set loop: the UNMODIFIED sample/optimize/sample_optimize_levelset.cpp is compiled and run
    (stdout history at 6 digits + its last result*.vtk), and the same loop is run through the reference's routines by
    oracle/ref_shim.cpp ref_levelset_run for a full-precision history."""
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(f"{tmp}/sample/optimize")
        exe = os.path.join(tmp, "levelset")
        subprocess.run(["g++", "-O3", "-fopenmp", "-std=c++17", f"-I{REF}", f"{REF}/sample/optimize/sample_optimize_levelset.cpp", "-o", exe], check=True)
        txt = subprocess.run([exe], check=True, capture_output=True, text=True, cwd=tmp).stdout
        rows = [[float(v) for v in re.findall(r"= ([-+0-9.e]+)", ln)] for ln in txt.split("\n") if ln.startswith("t =")]
        out["stdout_hist"] = np.array(rows)                      # t, compliance (= objective/nelem), volume, lambda
        out["stdout_converged"] = np.int64("Convergence" in txt)
        files = sorted([f for f in os.listdir(f"{tmp}/sample/optimize") if f.startswith("result")], key=lambda f: int(f[6:-4]))
        f = parse_vtk_fields(f"{tmp}/sample/optimize/{files[-1]}")
        out["vtk_count"] = np.int64(len(files))
        out["vtk_u"], out["vtk_phi"], out["vtk_str"] = f["u"][:, :2], f["phi"], f["str"]
    reflib.set_num_threads(8)
    P = problems.levelset2d()
    R = reflib.levelset_run(P.coords, P.conn, P.fixed, P.loads, P.phifixed, P.prm(), P.tmax, np.ones(P.nnode), np.ones(P.nelem))
    out["hist"], out["phi"], out["str"], out["u"] = R["hist"], R["phi"], R["str"], R["u"]
    out["iters"], out["converged"] = np.int64(R["iters"]), np.int64(R["converged"])
    R40 = reflib.levelset_run(P.coords, P.conn, P.fixed, P.loads, P.phifixed, P.prm(), 40, np.ones(P.nnode), np.ones(P.nelem))
    out["phi40"], out["str40"] = R40["phi"], R40["str"]
    # a second, smaller case started from a perturbed state (holes), 25 iterations
    P2 = problems.levelset2d(24, 16, nvol=10.0)
    rng = np.random.default_rng(3)
    phi0 = np.clip(rng.uniform(-0.3, 1.0, P2.nnode), -1, 1)
    str0 = (phi0[P2.conn].mean(axis=1) >= 0).astype(float)
    R2 = reflib.levelset_run(P2.coords, P2.conn, P2.fixed, P2.loads, P2.phifixed, P2.prm(), 25, phi0, str0)
    out.update(small_phi0=phi0, small_str0=str0, small_hist=R2["hist"], small_phi=R2["phi"], small_str=R2["str"], small_iters=np.int64(R2["iters"]))
    np.savez_compressed(f"{OUT}/levelset.npz", **out)
    print("levelset", out["stdout_hist"].shape, int(out["vtk_count"]), int(out["iters"]), bool(out["converged"]), R2["hist"][-1])


def krylov_case():
    """A non-symmetric test system: the Q4 conduction matrix of heat2d(14, 10) plus a skew, convection-like perturbation on the
    same pattern (deterministic)."""
    from oracle import portlib as orc
    P = problems.heat2d(14, 10)
    S = orc.assemble(P.eq, P.coords, P.conn, P.fixed, P.loads, np.ones(P.nelem))[0]
    indptr, indices, data, F = S.arrays()
    rng = np.random.default_rng(5)
    rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    data2 = data + 0.3 * np.sign(indices - rows) * rng.uniform(0.5, 1.0, len(data)) * np.abs(data)
    b = rng.uniform(-1, 1, len(F))
    return indptr, indices, data2, b


def golden_krylov():
    """BiCGSTAB / BiCGSTAB2 / ScalingBiCGSTAB / ILU0BiCGSTAB (CG.h:159-253, 357-393, 458-495) of the live reference."""
    reflib.set_num_threads(1)
    indptr, indices, data, b = krylov_case()
    A = reflib.system_from_csr(indptr, indices, data)
    d = dict(indptr=indptr, indices=indices, data=data, b=b)
    for kind, nm in ((3, "bicgstab"), (4, "bicgstab2"), (5, "scalingbicgstab"), (6, "ilu0bicgstab")):
        d[f"x_{nm}"] = A.solve(kind, b)[0]
    d["x_exact"] = np.linalg.solve(_dense(indptr, indices, data), b)
    # the reference's own non-symmetric sample: sample/advection/sample_advectiondiffusion_static.cpp (T3, SUPG), solved there with
    # BiCGSTAB; its committed output is AdvectionSUPG.vtk.  K and F are assembled by the reference's routines (ref_advection_system).
    adv = f"{REF}/sample/advection"
    nodes = np.array([[float(v) for v in r[1:3]] for r in csv_rows(f"{adv}/Node.csv")])
    elems = np.array([[int(v) for v in r[1:4]] for r in csv_rows(f"{adv}/Element.csv")], dtype=np.int32)
    fix = [(int(r[0]), float(r[1])) for r in csv_rows(f"{adv}/Dirichlet.csv") if r[1] != "free"]
    fn, fv = np.array([f[0] for f in fix], np.int32), np.array([f[1] for f in fix])
    S = reflib.advection_system(nodes, elems, fn, fv)
    ai, aj, ad, aF = S.arrays()
    L = open(f"{adv}/AdvectionSUPG.vtk").read().split("\n")
    i0 = next(k for k, ln in enumerate(L) if ln.startswith("SCALARS T"))
    T = np.array([float(v) for v in L[i0 + 2:i0 + 2 + len(nodes)]])
    d.update(adv_coords=nodes, adv_conn=elems, adv_fix_node=fn, adv_fix_val=fv, adv_indptr=ai, adv_indices=aj, adv_data=ad, adv_F=aF,
             adv_n2g=S.nodetoglobal(len(nodes), 1), adv_T=T, adv_x_bicgstab=S.solve(3, aF)[0])
    np.savez_compressed(f"{OUT}/live_krylov.npz", **d)
    print("krylov", len(b), [float(np.abs(d[k] - d["x_exact"]).max()) for k in d if k.startswith("x_") and k != "x_exact"])


def advection_sample_inputs():
    """Node.csv / Element.csv / Dirichlet.csv / DirichletD.csv of sample/advection, the dynamic sample's initial cone
    (sample_advectiondiffusion_dynamic.cpp:27-34) and its per-element rotating velocity (:48-50, CenterOfGravity General.h:71-78)."""
    adv = f"{REF}/sample/advection"
    nodes = np.array([[float(v) for v in r[1:3]] for r in csv_rows(f"{adv}/Node.csv")])
    elems = np.array([[int(v) for v in r[1:4]] for r in csv_rows(f"{adv}/Element.csv")], dtype=np.int32)
    fix = [(int(r[0]), float(r[1])) for r in csv_rows(f"{adv}/Dirichlet.csv") if r[1] != "free"]
    fixd = [(int(r[0]), float(r[1])) for r in csv_rows(f"{adv}/DirichletD.csv") if r[1] != "free"]
    T0 = np.zeros(len(nodes))
    r = np.sqrt((nodes[:, 0] - 0.5) ** 2 + (nodes[:, 1] - 0.75) ** 2)
    T0[r <= 0.25] = 0.5 * (np.cos(4.0 * np.pi * r[r <= 0.25]) + 1.0)
    cg = (nodes[elems[:, 0]] + nodes[elems[:, 1]] + nodes[elems[:, 2]]) / 3.0
    vel = np.stack([-(cg[:, 1] - 0.5), cg[:, 0] - 0.5], axis=1)
    return dict(coords=nodes, conn=elems, fix_node=np.array([f[0] for f in fix], np.int32), fix_val=np.array([f[1] for f in fix]),
                fixd_node=np.array([f[0] for f in fixd], np.int32), fixd_val=np.array([f[1] for f in fixd]), T0=T0, vel=vel)


def vtk_scalar(path, name, n):
    L = open(path).read().split("\n")
    i0 = next(k for k, ln in enumerate(L) if ln.startswith(f"SCALARS {name}"))
    return np.array([float(v) for v in L[i0 + 2:i0 + 2 + n]])


def golden_advection():
    """Advection-diffusion element family (Advection.h:19-229) of the live reference: every routine on
    every 2-D <SF, IC> on a distorted element, the systems of the two advection samples, a Q4 / T6 / Q8 time step with all six routines,
    and the committed outputs AdvectionSUPG.vtk, AdvectionSUPGdynamic0.vtk, AdvectionSUPGdynamic99.vtk."""
    from pansfem2_b200 import eqcode as ec, mesher
    reflib.set_num_threads(1)
    rng = np.random.default_rng(20211003)
    d = {}
    cases = []
    for shape in (ec.SHAPE_T3, ec.SHAPE_T6, ec.SHAPE_Q4, ec.SHAPE_Q8):
        nat = mesher.NATURAL_NODES[ec.SHAPE_NAME[shape]]
        for quad in ec.SHAPE_RULES[shape]:
            xe = nat * np.array([1.3, 0.9]) + 0.08 * rng.uniform(-1, 1, nat.shape)
            for terms in (1, 2, 4, 8, 16, 32, 7, 55, 63):
                for k in (1.0e-6, 0.4):           # alpha > 3 and alpha <= 3 (Advection.h:78-82)
                    ax, ay = rng.uniform(-1.5, 1.5, 2)
                    cases.append((shape, quad, terms, ax, ay, k))
                    d[f"xe_{len(cases) - 1}"] = xe
                    d[f"ke_{len(cases) - 1}"] = reflib.advdiff_element(shape, quad, terms, xe, ax, ay, k)
    d["cases"] = np.array(cases)
    # the dynamic sample: first step's system, the field after steps 1, 2, 3 and 100 (live reference, BiCGSTAB as there), committed VTKs
    I = advection_sample_inputs()
    d.update({f"smp_{k}": v for k, v in I.items()})
    adv = f"{REF}/sample/advection"
    n = len(I["coords"])
    d["smp_T_static_vtk"] = vtk_scalar(f"{adv}/AdvectionSUPG.vtk", "T", n)
    d["smp_T_dyn0_vtk"] = vtk_scalar(f"{adv}/AdvectionSUPGdynamic0.vtk", "T", n)
    d["smp_T_dyn99_vtk"] = vtk_scalar(f"{adv}/AdvectionSUPGdynamic99.vtk", "T", n)
    terms, dt, theta = 1 | 2 | 4 | 16 | 32, np.pi / 50.0, 0.5
    T = I["T0"].copy()
    T[I["fixd_node"]] = I["fixd_val"]
    free = None
    for step in range(100):
        S = reflib.advdiff_system(ec.SHAPE_T3, ec.QUAD_G1TRI, terms, I["coords"], I["conn"], I["fixd_node"], I["fixd_val"], I["vel"], 0.0, dt, theta, T)
        ai, aj, ad, aF = S.arrays()
        if step == 0:
            d.update(dyn_indptr=ai, dyn_indices=aj, dyn_data=ad, dyn_F=aF, dyn_n2g=S.nodetoglobal(n, 1))
            free = np.nonzero(d["dyn_n2g"][:, 0] >= 0)[0]
        x = S.solve(3, aF)[0]
        T[free] = x[d["dyn_n2g"][free, 0]]
        if step in (0, 1, 2, 99):
            d[f"dyn_T{step}"] = T.copy()
    print("dynamic sample vs committed VTKs:", np.abs(d["dyn_T0"] - d["smp_T_dyn0_vtk"]).max(), np.abs(d["dyn_T99"] - d["smp_T_dyn99_vtk"]).max())
    # one time step with all six routines on Q4 / T6 / Q8 family meshes, non-zero Dirichlet values, non-uniform velocity
    for nm, shape, quad, nn in (("q4", ec.SHAPE_Q4, ec.QUAD_G4SQ, (7, 5)), ("t6", ec.SHAPE_T6, ec.QUAD_G3TRI, (5, 4)), ("q8", ec.SHAPE_Q8, ec.QUAD_G9SQ, (4, 3))):
        coords, conn = mesher.family_mesh(ec.SHAPE_NAME[shape], nn)
        coords = coords + 0.07 * np.sin(1.7 * coords[:, ::-1] + 0.3)          # smooth distortion, keeps mid-side nodes consistent enough
        fn = np.nonzero(np.abs(coords[:, 0] - coords[:, 0].min()) < 0.2)[0].astype(np.int32)
        fv = 0.5 + 0.1 * np.arange(len(fn))
        cg = coords[conn].mean(axis=1)
        vel = np.stack([1.0 + 0.2 * cg[:, 1], -0.4 + 0.1 * cg[:, 0]], axis=1)
        Tn = np.cos(0.9 * coords[:, 0]) * np.sin(0.7 * coords[:, 1] + 0.2)
        S = reflib.advdiff_system(shape, quad, 63, coords, conn, fn, fv, vel, 0.02, 0.05, 0.6, Tn)
        ai, aj, ad, aF = S.arrays()
        d.update({f"{nm}_shape": np.int64(shape), f"{nm}_quad": np.int64(quad), f"{nm}_coords": coords, f"{nm}_conn": conn, f"{nm}_fix_node": fn,
                  f"{nm}_fix_val": fv, f"{nm}_vel": vel, f"{nm}_Tn": Tn, f"{nm}_indptr": ai, f"{nm}_indices": aj, f"{nm}_data": ad, f"{nm}_F": aF,
                  f"{nm}_x": S.solve(3, aF)[0]})
        print(nm, len(coords), "nodes", S.rows, "rows")
    # the theta-scheme heat conduction of sample/heattransfer/sample_heattransfer_dynamic.cpp (HeatTransfer + HeatCapacity, 500 steps) is the
    # Diffusion + Mass pair of the same family; its committed dynamic.vtk (the last step, on the mesh static.vtk also uses) is a golden too
    hv = f"{REF}/sample/heattransfer/dynamic.vtk"
    pts, cells = parse_vtk_mesh(hv)
    d["heatdyn_coords"], d["heatdyn_conn"] = pts[:, :2], cells
    d["heatdyn_T_vtk"] = vtk_scalar(hv, "T", len(pts))
    np.savez_compressed(f"{OUT}/live_advection.npz", **d)
    print("advection:", len(cases), "element cases")


def golden_homogenization():
    """sample/homogenization: the periodic node pairs of its Periodic.csv, the committed result_microscopic.vtk (characteristic displacements
    chi0..2 of the 20 x 20 cell with a 10 x 10 hole) and the two matrices the UNMODIFIED sample prints (check integral, homogenised C)."""
    hom = f"{REF}/sample/homogenization"
    pairs = np.array([[int(v) for v in r[:2]] for r in csv_rows(f"{hom}/Periodic.csv")], np.int32)
    pts, cells = parse_vtk_mesh(f"{hom}/result_microscopic.vtk")
    f = parse_vtk_fields(f"{hom}/result_microscopic.vtk")
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(f"{tmp}/sample/homogenization")
        shutil.copyfile(f"{hom}/Periodic.csv", f"{tmp}/sample/homogenization/Periodic.csv")
        exe = os.path.join(tmp, "homog")
        subprocess.run(["g++", "-O2", "-fopenmp", "-w", f"{hom}/sample_homogenization.cpp", "-o", exe], check=True)
        out = subprocess.run([exe], cwd=tmp, check=True, capture_output=True, text=True).stdout
    vals = np.array([float(v) for v in out.split()]).reshape(2, 3, 3)
    # sample/optimize/sample_optimize_homogenization.cpp: its Periodic.csv and the objective / weight history the UNMODIFIED driver prints
    # (157 iterations, about 6 minutes on 8 cores; the committed Homogenization_*.vtk are not reproduced by the current driver)
    opt = f"{REF}/sample/optimize"
    pairs_opt = np.array([[int(v) for v in r[:2]] for r in csv_rows(f"{opt}/Periodic.csv")], np.int32)
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(f"{tmp}/sample/optimize")
        shutil.copyfile(f"{opt}/Periodic.csv", f"{tmp}/sample/optimize/Periodic.csv")
        exe = os.path.join(tmp, "homopt")
        subprocess.run(["g++", "-O3", "-fopenmp", "-w", f"{opt}/sample_optimize_homogenization.cpp", "-o", exe], check=True)
        log = subprocess.run([exe], cwd=tmp, check=True, capture_output=True, text=True).stdout
    hist = np.array([[float(a), float(b)] for a, b in re.findall(r"Objective:\s*([-0-9.e+]+)\s*Weight:\s*([-0-9.e+]+)", log)])
    np.savez_compressed(f"{OUT}/homogenization.npz", pairs=pairs, coords=pts[:, :2], conn=cells, chi0=f["chi0"][:, :2], chi1=f["chi1"][:, :2],
                        chi2=f["chi2"][:, :2], check=vals[0], CH=vals[1], pairs_opt=pairs_opt, opt_history=hist)
    print("homogenization:", pairs.shape, pts.shape, cells.shape, vals[1])


def golden_plane_d():
    """PlaneStiffness / PlaneStiffnessBbar / PlaneStiffnessWilsonTaylor (Homogenization.h:141-280) of the live reference on distorted
    elements with symmetric and non-symmetric constitutive matrices."""
    from pansfem2_b200 import eqcode as ec, mesher
    rng = np.random.default_rng(20210120)
    d, cases = {}, []
    for shape in (ec.SHAPE_T3, ec.SHAPE_T6, ec.SHAPE_Q4, ec.SHAPE_Q8):
        nat = mesher.NATURAL_NODES[ec.SHAPE_NAME[shape]]
        for quad in ec.SHAPE_RULES[shape]:
            xe = nat * np.array([1.3, 0.9]) + 0.08 * rng.uniform(-1, 1, nat.shape)
            A = rng.uniform(-1, 1, (3, 3))
            Dsym = A @ A.T + 0.5 * np.eye(3)
            for mode, phys in ((0, ec.PHYS_PLANE_D), (1, ec.PHYS_PLANE_D_BBAR), (2, ec.PHYS_PLANE_D_WT)):
                if mode == 2 and (shape in (ec.SHAPE_T3, ec.SHAPE_T6) or quad == ec.QUAD_G1SQ):
                    continue
                for quad2 in (ec.SHAPE_RULES[shape] if mode == 1 else (0,)):
                    for D in (Dsym, Dsym + 0.1 * rng.uniform(-1, 1, (3, 3))):
                        k = len(cases)
                        cases.append(ec.eq_code(phys, shape, quad, quad2))
                        d[f"xe_{k}"], d[f"D_{k}"] = xe, D
                        d[f"ke_{k}"] = reflib.plane_d_element(shape, quad, quad2, mode, xe, D, 0.7)
    d["cases"] = np.array(cases, np.int64)
    np.savez_compressed(f"{OUT}/live_plane_d.npz", **d)
    print("plane_d:", len(cases), "cases")


def loadvec_cases():
    """(kind, shape, quad) of every load-vector routine the mirror exposes: surface routines on the two line shapes x two line rules,
    body routines on the four area shapes x the rules of their domain; kind 0 plane strain, 1 plane stress, 2 heat flux."""
    from pansfem2_b200 import eqcode as ec
    cases = []
    for kind in (0, 1, 2):
        for shape in (8, 9):
            for quad in (9, 10):
                cases.append((kind, shape, quad))
    for kind in (0, 1):
        for shape in (ec.SHAPE_T3, ec.SHAPE_T6, ec.SHAPE_Q4, ec.SHAPE_Q8):
            for quad in ec.SHAPE_RULES[shape]:
                cases.append((kind, shape, quad))
    return cases


LINE_NODES = {8: np.array([[0.0, 0.0], [1.0, 0.0]]), 9: np.array([[0.0, 0.0], [1.0, 0.0], [0.5, 0.0]])}


def golden_loadvec():
    """PlaneStrain / PlaneStress SurfaceForce + BodyForce and HeatTransferSurfaceFlux of the live reference (through their functor interface,
    affine force density) on distorted elements."""
    from pansfem2_b200 import eqcode as ec, mesher
    rng = np.random.default_rng(20211104)
    d, cases = {}, loadvec_cases()
    for k, (kind, shape, quad) in enumerate(cases):
        if shape in LINE_NODES:
            th = rng.uniform(0, 2 * np.pi)
            R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
            xe = (LINE_NODES[shape] * 1.7) @ R.T + rng.uniform(-1, 1, 2)
            if shape == 9:
                xe[2] += 0.1 * rng.uniform(-1, 1, 2)           # curved edge
        else:
            nat = mesher.NATURAL_NODES[ec.SHAPE_NAME[shape]]
            xe = nat * np.array([1.3, 0.9]) + 0.08 * rng.uniform(-1, 1, nat.shape)
        coef = rng.uniform(-2, 2, 6)
        d[f"xe_{k}"], d[f"coef_{k}"] = xe, coef
        d[f"fe_{k}"] = reflib.load_vector(kind, shape, quad, xe, coef, 0.7)
    d["cases"] = np.array(cases, np.int64)
    np.savez_compressed(f"{OUT}/live_loadvec.npz", **d)
    print("loadvec:", len(cases), "cases")


def golden_linalg():
    """Boundary types (Vector, Matrix, LILCSR, CSR host behaviour): tests/cpp/linalg_tables.cpp built against the reference's headers."""
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "linalg_tables")
        subprocess.run(["g++", "-O1", "-std=c++17", "-w", f"-I{REF}/src", f"{ROOT}/tests/cpp/linalg_tables.cpp", "-o", exe], check=True)
        open(f"{OUT}/linalg_tables.txt", "w").write(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)
    print("linalg golden written")


def golden_routines():
    """Host-side helpers without a device path (General.h, HeatTransferSurfaceFlux, ShapeFunction3Line, SetPeriodic ...):
    tests/cpp/host_routines.cpp built against the reference's headers."""
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "host_routines")
        subprocess.run(["g++", "-O1", "-std=c++17", "-w", f"-I{REF}/src", f"{ROOT}/tests/cpp/host_routines.cpp", "-o", exe], check=True)
        open(f"{OUT}/host_routines.txt", "w").write(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)
    print("host routines golden written")


def golden_meshers():
    """SquareMesh<T> / SquareMesh2<T> of the reference (PrePost/Mesher/SquareMesh.h): tests/cpp/mesher_tables.cpp built against the
    reference's headers; the mirror build must print the same table."""
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "mesher_tables")
        subprocess.run(["g++", "-O1", "-std=c++17", "-w", f"-I{REF}/src", f"{ROOT}/tests/cpp/mesher_tables.cpp", "-o", exe], check=True)
        open(f"{OUT}/mesher_tables.txt", "w").write(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)
    print("mesher golden written")


def golden_io():
    """The data formats either side of the path: tests/cpp/io_formats.cpp built against the REFERENCE's PrePost headers (ExportToVTK.h,
    ImportFromCSV.h, ImportFromVTK.h / ImportFromVTK2.h); its VTK bytes and parsed values are the golden the mirror build must equal."""
    with tempfile.TemporaryDirectory() as tmp:
        for tag, flags in (("", []), ("_vtk2", ["-DIO_VTK2"])):
            exe = os.path.join(tmp, "io" + tag)
            subprocess.run(["g++", "-O1", "-std=c++17", "-w", *flags, f"-I{REF}/src", f"{ROOT}/tests/cpp/io_formats.cpp", "-o", exe], check=True)
            os.makedirs(os.path.join(tmp, "work"), exist_ok=True)
            out = subprocess.run([exe, "work"], cwd=tmp, check=True, capture_output=True, text=True).stdout
            open(f"{OUT}/io_formats{tag}.txt", "w").write(out)
        shutil.copyfile(os.path.join(tmp, "work", "out.vtk"), f"{OUT}/io_formats_out.vtk")
    print("io goldens written")


def _dense(indptr, indices, data):
    n = len(indptr) - 1
    M = np.zeros((n, n))
    for i in range(n):
        M[i, indices[indptr[i]:indptr[i + 1]]] = data[indptr[i]:indptr[i + 1]]
    return M


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "loadvec":
        golden_loadvec()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "krylov":
        golden_krylov()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "homogenization":
        golden_homogenization()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "plane_d":
        golden_plane_d()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "linalg":
        golden_linalg()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "routines":
        golden_routines()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "meshers":
        golden_meshers()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "io":
        golden_io()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "advection":
        golden_advection()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "levelset":
        golden_levelset()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "families":
        golden_families()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "conlin":
        f = parse_vtk_fields(f"{REF}/sample/optimize/Density_CONLIN.vtk")
        np.savez_compressed(f"{OUT}/density_conlin.npz", u=f["u"][:, :2], r=f["r"][:, :2], rho=f["s"])
        golden_conlin()
        sys.exit(0)
    golden_simp()
    golden_solid()
    golden_mma_kat()
    golden_live()
    golden_conlin()
    golden_families()
    golden_levelset()
    golden_krylov()
