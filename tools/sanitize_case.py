"""Small cases for compute-sanitizer (memcheck / racecheck; one tool per run, see profiles/r02_sanitizer.md):
  python tools/sanitize_case.py            one GPU: low-rank term (interacting on the coarse levels), 2 chains per launch,
                                           merged level-0 launches, graph replays + a cycle through the host API
  torchrun --nproc-per-node 2 tools/sanitize_case.py strips     two ranks: row strips of one lattice
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multigridmc_b200 as m  # noqa: E402
from multigridmc_b200 import workloads as w  # noqa: E402


def single():
    n, nlevel = 256, 5
    loc, _, _, var = w.measurement_set(8)
    B = w.point_measurement_matrix(n, n, loc, var, 1e-3)
    ctx = m.Context(n, n, nlevel, B=B, seed=7, nchains=2)
    rng = np.random.default_rng(0)
    nd = ctx.ndof()
    ctx.set_rhs(rng.standard_normal(2 * nd))
    ctx.set_state(rng.standard_normal(2 * nd))
    ctx.set_qoi([nd // 2], [1.0])
    ctx.set_philox_position(0)
    z = ctx.sample(4)
    ctx.sample(6, series=False)  # (more cycles than the series buffer of the earlier call holds)
    x = ctx.get_state()
    assert np.all(np.isfinite(x)) and np.all(np.isfinite(z))
    # deterministic twin: V(2,2) multigrid-preconditioned Richardson on the same hierarchy
    c2 = m.Context(n, n, nlevel, B=B, npresmooth=2, npostsmooth=2)
    xs, hist, it, cv = c2.loop_solve(rng.standard_normal(nd), rtol=1e-10, atol=1e-15, maxiter=20)
    assert hist[-1] < 1e-2 * hist[0], (hist[0], hist[-1])
    print("sanitize_case single ok", float(z[-1, 0]), len(hist))


def strips():
    import torch
    import torch.distributed as dist
    from multigridmc_b200 import strips as st

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    n, nlevel = 1024, 6
    loc, _, _, var = w.measurement_set(4)
    B = w.point_measurement_matrix(n, n, loc, var, 1e-3)
    dev = int(os.environ.get("LOCAL_RANK", rank))
    ctx = m.Context(n, n, nlevel, B=B, seed=7, device=dev, strip_rank=rank, strip_nranks=world)
    st.connect(ctx, dist, torch.device("cuda", dev))
    rng = np.random.default_rng(0)
    nd = ctx.ndof()
    ctx.set_rhs(rng.standard_normal(nd))
    ctx.set_state(rng.standard_normal(nd))
    ctx.set_philox_position(0)
    ctx.sample(3, series=False)
    x = ctx.get_state()
    assert np.all(np.isfinite(x))
    dist.barrier()
    print("sanitize_case strips ok", rank)
    dist.destroy_process_group()


if __name__ == "__main__":
    strips() if len(sys.argv) > 1 and sys.argv[1] == "strips" else single()
