"""Development aid (2 GPUs): radius-2 row strips vs single GPU, error by row."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def worker(rank, world, port, n, nlevel, nsamples, minsites, npre=1, npost=1):
    if minsites: os.environ["MGMC_STRIP_MIN_SITES"] = str(minsites)
    import torch, torch.distributed as dist
    import multigridmc_b200 as m
    from multigridmc_b200 import strips
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    rng = np.random.default_rng(5)
    nd = (n - 1) ** 2
    f, x0 = rng.standard_normal(nd), rng.standard_normal(nd)
    pde = "squared_shiftedlaplace_fd"
    def run(ctx):
        ctx.set_rhs(f); ctx.set_state(x0); ctx.set_qoi([nd // 2], [1.0]); ctx.set_philox_position(0)
        dist.barrier()
        ctx.sample(nsamples)
        return ctx.get_state()
    ref = m.Context(n, n, nlevel, device=rank, seed=99, pde=pde, npresmooth=npre, npostsmooth=npost)
    x_ref = run(ref)
    ctx = m.Context(n, n, nlevel, device=rank, seed=99, strip_rank=rank, strip_nranks=world, pde=pde, npresmooth=npre, npostsmooth=npost)
    strips.connect(ctx, dist, dev)
    x_loc = run(ctx)
    x = strips.gather_state(x_loc, ctx.desc, dist, dev)
    if rank == 0:
        e = np.abs(x - x_ref).reshape(n - 1, n - 1).max(axis=1)
        bad = np.nonzero(e > 0)[0]
        print(f"n={n} L={nlevel} samples={nsamples} minsites={minsites} npre={npre} npost={npost}: scale {np.abs(x_ref).max():.3e} max err {e.max():.3e} bad rows {len(bad)}", bad[:10] + 1, bad[-10:] + 1, "err flag", ctx.strip_error(), flush=True)
        for lvl in range(nlevel):
            print("   level", lvl, m.strip_partition(ctx.desc, lvl, 0), m.strip_partition(ctx.desc, lvl, 1))
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    port = 29700
    for (n, nlevel, ns, ms, npre, npost) in [(256, 3, 1, 1, 0, 0), (256, 4, 3, 1, 1, 1)]:
        port += 1
        mp.spawn(worker, args=(2, port, n, nlevel, ns, ms, npre, npost), nprocs=2, join=True)
