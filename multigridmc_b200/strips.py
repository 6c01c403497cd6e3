"""Row-strip decomposition of one lattice over the GPUs of a node: the host-side plumbing around
``mgmc_strip_*`` (include/mgmc_b200.h).  One process per GPU; ``torch.distributed`` only carries the
64-byte CUDA IPC handles at set-up and the scalar QoI series at the end -- the halo rows themselves
move inside the kernels' own stream (peer stores over NVLink + device-side flags), never through
NCCL or the host."""
import numpy as np


def exchange_blobs(blob, dist, device=None):
    """all-gather one small bytes object per rank (works with the nccl and the gloo backend)."""
    import torch

    world = dist.get_world_size()
    t = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [bytes(o.cpu().numpy().tobytes()) for o in out]


def connect(ctx, dist, device):
    """Map the neighbours' arenas into this rank's context (collective over all ranks)."""
    ctx.strip_connect(exchange_blobs(ctx.strip_export(), dist, device))
    dist.barrier()


def reduce_series(series, dist, device=None):
    """QoI series: every rank holds the partial dot product over the rows it owns; sum over ranks."""
    import torch

    t = torch.as_tensor(np.ascontiguousarray(series), dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def own_rows(desc, level, rank):
    """(row_lo, row_hi) owned by `rank` on `level` (all rows on a replicated level)."""
    from . import capi

    lo, hi, _ = capi.strip_partition(desc, level, rank)
    return lo, hi


def own_slice(desc, rank, level=0):
    """slice of the lexicographic interior vector that holds the rows owned by `rank`."""
    lo, hi = own_rows(desc, level, rank)
    w = (desc.nx >> level) - 1
    return slice((lo - 1) * w, hi * w)


def gather_state(x_local, desc, dist, device=None):
    """Assemble the global state from the ranks' own rows (every rank passes its full-size local vector)."""
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    mine = np.zeros_like(x_local)
    sl = own_slice(desc, rank)
    mine[sl] = x_local[sl]
    t = torch.as_tensor(mine)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
