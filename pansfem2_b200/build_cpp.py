"""Build the C++ side of the drop-in boundary into pansfem2_b200/bin (git-ignored; travels to the GPU box):

  * our own drivers under pansfem2_b200/sample (batched, device-resident API);
  * where /root/reference exists: the UNMODIFIED reference drivers (sample_optimize_density_oc.cpp, ..._mma.cpp, ..._CONLIN.cpp,
    sample/solid/sample_linear.cpp) compiled against the header mirror pansfem2_b200/src instead of the reference's src/ -
    the drop-in check.  The reference sources are only symlinked into a scratch tree, never copied into the repo.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "bin")
REFERENCE = "/root/reference"
CXX = "g++"
FLAGS = ["-O2", "-std=c++17", "-fopenmp"]
LINK = ["-L" + HERE, "-lpansfem2_b200", "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/.."]

OWN = [("sample/optimize/sample_optimize_density_batched.cpp", "sample_optimize_density_batched"),
       ("sample/optimize/sample_optimize_density_families.cpp", "sample_optimize_density_families"),
       ("sample/optimize/sample_optimize_levelset_batched.cpp", "sample_optimize_levelset_batched"),
       ("sample/advection/sample_advectiondiffusion_batched.cpp", "sample_advectiondiffusion_batched"),
       ("sample/planestrain/sample_planestrain_batched.cpp", "sample_planestrain_batched")]
DROPIN = [("sample/optimize/sample_optimize_density_oc.cpp", "dropin_density_oc"),
          ("sample/optimize/sample_optimize_density_mma.cpp", "dropin_density_mma"),
          ("sample/optimize/sample_optimize_density_CONLIN.cpp", "dropin_density_conlin"),
          ("sample/solid/sample_linear.cpp", "dropin_solid_linear"),
          ("sample/planestrain/sample_planestrain.cpp", "dropin_planestrain_t3"),
          ("sample/optimize/sample_optimize_levelset.cpp", "dropin_levelset"),
          ("sample/advection/sample_advectiondiffusion_static.cpp", "dropin_advection_static"),
          ("sample/advection/sample_advectiondiffusion_dynamic.cpp", "dropin_advection_dynamic"),
          ("sample/homogenization/sample_homogenization.cpp", "dropin_homogenization"),
          ("sample/optimize/sample_optimize_homogenization.cpp", "dropin_optimize_homogenization")]


def _compile(src, exe):
    if os.path.exists(exe) and os.path.getmtime(exe) > max(os.path.getmtime(src), _hdr_mtime()):
        return exe
    r = subprocess.run([CXX, *FLAGS, src, "-o", exe, *LINK], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{CXX} failed on {src}:\n{r.stderr[-4000:]}")
    return exe


def _hdr_mtime():
    m = os.path.getmtime(os.path.join(HERE, "libpansfem2_b200.so"))
    for dp, _, files in os.walk(os.path.join(HERE, "src")):
        for f in files:
            m = max(m, os.path.getmtime(os.path.join(dp, f)))
    return m


def build_all():
    os.makedirs(BIN, exist_ok=True)
    built = []
    for rel, name in OWN:
        built.append(_compile(os.path.join(HERE, rel), os.path.join(BIN, name)))
    if os.path.isdir(os.path.join(REFERENCE, "sample")):
        scratch = os.path.join(HERE, "csrc", "build", "dropin")
        shutil.rmtree(scratch, ignore_errors=True)
        os.makedirs(scratch)
        os.symlink(os.path.join(HERE, "src"), os.path.join(scratch, "src"))
        for rel, name in DROPIN:
            dst = os.path.join(scratch, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            os.symlink(os.path.join(REFERENCE, rel), dst)
            built.append(_compile(dst, os.path.join(BIN, name)))
    return built


if __name__ == "__main__":
    for b in build_all():
        print(b)
