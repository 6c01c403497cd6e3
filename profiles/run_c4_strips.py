#!/usr/bin/env python
"""BASELINE config C4: driver_mgmc, 2-d squared shifted Laplacian (biharmonic type) 2048 x 2048, 7 levels, V(1,1) SSOR,
prior, ONE chain domain-decomposed into row strips over the ranks (per-colour halo exchange inside the colour launches,
NVLink peer stores).  Launch with torchrun, one rank per GPU; rank 0 prints one JSON line.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/run_c4_strips.py [--n 2048] [--nlevel 7]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--n", type=int, default=2048)
    p.add_argument("--nlevel", type=int, default=7)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=10)
    a = p.parse_args()
    import torch
    import torch.distributed as dist

    import multigridmc_b200 as m
    from multigridmc_b200 import strips

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, nd = a.n, (a.n - 1) ** 2
    xs = np.arange(1, n) / n
    u = np.outer(np.sin(np.pi * xs), np.sin(np.pi * xs)).ravel()
    kw = dict(Lambda=0.2, pde="squared_shiftedlaplace_fd", device=local, seed=5418513)
    ref = m.Context(n, n, a.nlevel, **kw)
    f = ref.op_apply(0, u)
    ref.close()
    if world > 1:
        ctx = m.Context(n, n, a.nlevel, strip_rank=rank, strip_nranks=world, **kw)
        strips.connect(ctx, dist, torch.device("cuda", local))
    else:
        ctx = m.Context(n, n, a.nlevel, **kw)
    ctx.set_rhs(f)
    ctx.set_state(np.zeros(nd))
    ctx.set_qoi([nd // 2], [1.0])
    ctx.set_philox_position(0)
    if world > 1:
        dist.barrier()
    ctx.sample(a.warmup, series=False)
    if world > 1:
        dist.barrier()
    ms, _ = ctx.sample_timed(a.steps, series=False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    byts, upd = ctx.cycle_model()
    err = int(ctx.strip_error()) if world > 1 else 0
    dist_levels = sum(1 for l in range(a.nlevel) if m.strip_partition(ctx.desc, l, 0)[2]) if world > 1 else 0
    if rank == 0:
        sec = ms / a.steps * 1e-3
        print(json.dumps({"config": f"C4 squared_shiftedlaplace_fd {n}x{n} L{a.nlevel} V(1,1) SSOR, one chain on row strips", "n_gpus": world,
                          "distributed_levels": dist_levels, "ms_per_cycle": ms / a.steps, "samples_per_s": 1.0 / sec,
                          "site_updates_per_s": upd / sec, "error_flag": err}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
