"""GPU tests of the row-partitioned path (csrc/dist.cu, p2p.cuh, pcg_persistent.cuh DIST instantiations, pf2_simp_set_partition).

The reference has no distributed code (its only parallelism is the OpenMP loop of CSR.h:114); the partitioned loop must reproduce the
single-GPU loop.  With >= 2 GPUs on the box, tests/dist_worker.py runs under torchrun on 2 ranks for OC and MMA, 2-D / heat / hex8, both
backends (peer memory, NCCL) and both PCG forms, and asserts objective, u and the design after 4 iterations against the single-GPU loop.
On a 1-GPU box those cases skip and the world-size-1 case below still drives the DIST instantiation (allreduce warp, epochs) of the kernels.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from pansfem2_b200 import capi, partition, problems

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


CASES = [
    # default: three-kernel loop over peer memory (LL allreduce in the last CTA, halo planes pushed by the p-update kernel)
    ("2d 96 40", {}), ("2d 96 40 --mma", {}), ("3d 16 8 6", {}), ("3d 16 8 6 --mma", {}), ("heat 48 48", {}),
    ("2d 96 40 --warm", {}), ("3d 16 8 6 --warm", {"PF2_P2P": "0"}),             # pf2_solve_x0 under the partition (ghost entries of x0 valid)
    ("2d 96 40", {"PF2_HALO_DEFER": "1"}), ("3d 16 8 6", {"PF2_HALO_DEFER": "1", "PF2_SELL_L2_MB": "100"}),      # opt-in variants
    # the persistent kernel's partitioned instantiation
    ("2d 96 40", {"PF2_PCG": "1"}), ("2d 96 40 --mma", {"PF2_PCG": "1"}), ("heat 48 48", {"PF2_PCG": "1"}), ("2d 96 40 --warm", {"PF2_PCG": "1"}),
    # NCCL backend
    ("2d 96 40 --mma", {"PF2_P2P": "0"}), ("3d 16 8 6", {"PF2_P2P": "0"}),
    ("3d 16 8 6 --matrix-free", {}),
    # ILU0CG under the partition: block-Jacobi ILU(0) per rank (CG.h:258-352), both backends
    ("2d 96 40 --ilu", {}), ("3d 16 8 6 --ilu", {"PF2_P2P": "0"}),
    # single-reduction PCG (pf2_csr_set_cg_variant / PF2_CG_SINGLE_REDUCTION): one cross-GPU sum, two kernels per iteration; same solution to the
    # solver tolerance
    ("2d 96 40", {"PF2_CG_SINGLE_REDUCTION": "1"}), ("3d 16 8 6 --mma", {"PF2_CG_SINGLE_REDUCTION": "1"}), ("heat 48 48", {"PF2_CG_SINGLE_REDUCTION": "1"}),
    ("2d 96 40 --warm", {"PF2_CG_SINGLE_REDUCTION": "1"}), ("3d 16 8 6 --matrix-free", {"PF2_CG_SINGLE_REDUCTION": "1"}),
]


@pytest.mark.parametrize("args,env", CASES)
def test_partitioned_loop_on_two_gpus_equals_single_gpu_loop(args, env):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs on the box (run with gpurun --gpus 2)")
    e = dict(os.environ, **env)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tests", "dist_worker.py"), *args.split(), "--iters", "4"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=e, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-1500:] + "\n".join(l for l in r.stderr.splitlines() if "Error" in l or "assert" in l or "pf2" in l)[-3000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert res["world"] == 2 and res["max_f_rel"] < 1e-8 and res["max_s_diff"] < 1e-6
    assert (res["pcg"]["single_reduction_solves"] >= 4) == (env.get("PF2_CG_SINGLE_REDUCTION") == "1"), res["pcg"]     # what ran is what was asked for
    if env.get("PF2_PCG") != "1":
        assert res["pcg"]["solves"] == 0
    elif args == "2d 96 40":
        assert res["pcg"]["solves"] > 0          # the persistent kernel's partitioned instantiation is what ran (every slab qualifies)


@pytest.mark.parametrize("make", [lambda: problems.cantilever2d(48, 32, opt_kind=problems.OPT_MMA, filter_kind=problems.FILTER_DENSITY),
                                  lambda: problems.cantilever3d(10, 6, 4)])
@pytest.mark.parametrize("pcg", ["0", "1", "ilu", "cg1"])
def test_partitioned_path_with_one_rank_equals_plain_loop(make, pcg, monkeypatch):
    """world size 1: no neighbours, but every reduction goes through the peer-memory LL allreduce of the partitioned kernels
    (pcg = 1: the persistent kernel's partitioned instantiation through pf2_csr_set_pcg_mode)."""
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29741")
        dist.init_process_group("gloo", rank=0, world_size=1)
    P = make()
    ctx = capi.Context(0)
    solver = capi.SOLVER_ILU0CG if pcg == "ilu" else capi.SOLVER_SCALINGCG
    ref = capi.Simp(ctx, P, solver=solver)
    fr = [ref.iterate(check_convergence=False) for _ in range(3)]
    o = ref.get()
    ref.close()
    D = capi.Dist(ctx, 0, 1)
    S = partition.slab(P, 0, 1)
    sim = capi.Simp(ctx, S.local, solver=solver)
    D.set_simp_partition(sim, S, P.nelem)
    sim.A.set_pcg_mode(1 if pcg == "1" else 0)
    sim.A.set_cg_variant(1 if pcg == "cg1" else 0)           # cg1: the single-reduction recurrences (two kernels, one sum per iteration)
    fd = [sim.iterate(check_convergence=False) for _ in range(3)]
    assert sim.A.pcg_stats()["single_reduction_solves"] == (3 if pcg == "cg1" else 0)
    od = sim.get()
    assert pcg == "1" or sim.A.pcg_stats()["solves"] == 0      # (ragged SELL-C-sigma slabs fall back to the three-kernel loop under pcg = 1)
    for a, b in zip(fd, fr):
        assert abs(a["f"] - b["f"]) < 1e-9 * abs(b["f"]) and abs(a["cg_iters"] - b["cg_iters"]) <= 3 + b["cg_iters"] // 50
    assert np.abs(od["s"] - o["s"]).max() < 1e-7 and np.abs(od["u"] - o["u"]).max() < 1e-8 * np.abs(o["u"]).max()
    sim.close()
    D.close()
    ctx.close()
