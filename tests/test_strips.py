"""Row-strip decomposition.

CPU part (gloo, world_size 2): partition covers every row exactly once and starts on tile boundaries,
handle exchange and QoI / state reductions of multigridmc_b200.strips.
GPU part (needs >= 2 GPUs, skipped on a single-GPU box; run with `gpurun --gpus 2`): the chain advanced
cooperatively by 2 ranks is BIT-IDENTICAL to the single-GPU chain, and so is the QoI series."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,nlevel,nranks", [(4096, 8, 2), (4096, 8, 8), (2048, 7, 4), (256, 4, 2), (1024, 6, 8)])
def test_partition_covers_every_row_once(n, nlevel, nranks):
    import multigridmc_b200 as m
    from multigridmc_b200 import capi

    os.environ["MGMC_STRIP_MIN_SITES"] = "1"
    desc = capi.make_desc(n, n, nlevel, strip_nranks=nranks)
    ndist = 0
    for level in range(nlevel):
        ny = n >> level
        parts = [m.strip_partition(desc, level, r) for r in range(nranks)]
        if parts[0][2]:
            ndist += 1
            rows = []
            for lo, hi, dist in parts:
                assert dist and (lo - 1) % 8 == 0  # strips start on tile boundaries
                rows += list(range(lo, hi + 1))
            assert rows == list(range(1, ny))
        else:
            assert all(p == (1, ny - 1, False) for p in parts)
    assert 1 <= ndist < nlevel
    with pytest.raises(m.MgmcError):
        m.strip_partition(capi.make_desc(64, 64, 3, strip_nranks=2), 0, 0)


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from multigridmc_b200 import capi, strips

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    blobs = strips.exchange_blobs(bytes([rank + 1] * 64), dist)
    ok = blobs == [bytes([r + 1] * 64) for r in range(world)]
    series = strips.reduce_series(np.full(5, rank + 1.0), dist)
    ok = ok and np.array_equal(series, np.full(5, sum(range(1, world + 1))))
    desc = capi.make_desc(256, 256, 4, strip_nranks=world)
    x = np.full(255 * 255, float(rank + 1))
    g = strips.gather_state(x, desc, dist)
    for r in range(world):
        ok = ok and np.all(g[strips.own_slice(desc, r)] == r + 1)
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


def test_host_plumbing_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == {0: True, 1: True}


def _gpu_worker(rank, world, port, q, n, nlevel, nsamples, nmeas, pde="shiftedlaplace_fd"):
    sys.path.insert(0, ROOT)
    os.environ["MGMC_STRIP_MIN_SITES"] = "1"  # small test lattices: distribute as many levels as the strips allow
    import torch
    import torch.distributed as dist

    import multigridmc_b200 as m
    from multigridmc_b200 import strips
    from multigridmc_b200 import workloads as w

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    rng = np.random.default_rng(5)
    nd = (n - 1) ** 2
    f, x0 = rng.standard_normal(nd), rng.standard_normal(nd)
    B = None
    if nmeas:
        loc, _, _, var = w.measurement_set(nmeas)
        B = w.point_measurement_matrix(n, n, loc, var, 1e-3)
    qidx = np.array([nd // 2 + 7, 3 * (n - 1) + 5, nd - 10])
    qval = np.array([1.0, -2.0, 0.5])

    def run(ctx):
        ctx.set_rhs(f)
        ctx.set_state(x0)
        ctx.set_qoi(qidx, qval)
        ctx.set_philox_position(0)
        dist.barrier()
        z = ctx.sample(nsamples)[:, 0]
        return ctx.get_state(), z

    ref = m.Context(n, n, nlevel, B=B, device=rank, seed=99, pde=pde)
    x_ref, z_ref = run(ref)
    ctx = m.Context(n, n, nlevel, B=B, device=rank, seed=99, strip_rank=rank, strip_nranks=world, pde=pde)
    strips.connect(ctx, dist, dev)
    x_loc, z_part = run(ctx)
    err = ctx.strip_error()
    z = strips.reduce_series(z_part, dist, dev)
    sl = strips.own_slice(ctx.desc, rank)
    x = strips.gather_state(x_loc, ctx.desc, dist, dev)
    # rows far away from the own strip were NOT computed here: they still hold the initial state
    other = strips.own_slice(ctx.desc, (rank + 1) % world)
    far = x_loc[other][(n - 1) * 40:-(n - 1) * 40]
    far0 = x0[other][(n - 1) * 40:-(n - 1) * 40]
    res = dict(rank=rank, err=err, own=float(np.abs(x_loc[sl] - x_ref[sl]).max()), glob=float(np.abs(x - x_ref).max()),
               series=float(np.abs(z - z_ref).max()), scale=float(np.abs(x_ref).max()), launches=int(ctx.launch_count()),
               ref_launches=int(ref.launch_count()), far_untouched=bool(np.array_equal(far, far0)))
    dist.barrier()
    dist.destroy_process_group()
    q.put(res)


def _run_gpu_case(n, nlevel, nsamples, nmeas, world=2, pde="shiftedlaplace_fd"):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_gpu_worker, args=(r, world, port, q, n, nlevel, nsamples, nmeas, pde)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
    return res


def _ngpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
@pytest.mark.parametrize("n,nlevel,nmeas", [(256, 4, 0), (1024, 6, 0), (1024, 6, 8)])
def test_two_rank_chain_is_bit_identical_to_single_gpu(n, nlevel, nmeas):
    for r in _run_gpu_case(n, nlevel, 5, nmeas):
        assert r["err"] == 0, r
        assert r["own"] == 0.0 and r["glob"] == 0.0 and r["series"] <= 1e-12 * max(r["scale"], 1.0), r
        assert r["far_untouched"] and r["launches"] > r["ref_launches"], r  # really decomposed: wait / push launches, untouched far rows


@pytest.mark.gpu
@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
@pytest.mark.parametrize("n,nlevel", [(256, 4), (512, 5)])
def test_two_rank_biharmonic_chain_is_bit_identical_to_single_gpu(n, nlevel):
    """BASELINE config 4: squared shifted Laplacian (radius-2 stencils, 9 colours) on row strips, halo rows exchanged
    per colour inside the colour launches over NVLink peer memory."""
    for r in _run_gpu_case(n, nlevel, 4, 0, pde="squared_shiftedlaplace_fd"):
        assert r["err"] == 0, r
        assert r["own"] == 0.0 and r["glob"] == 0.0 and r["series"] <= 1e-12 * max(r["scale"], 1.0), r
        assert r["far_untouched"], r
