"""k-range sharding of the eigenpair loop across GPUs (one process per GPU).

The reference's only parallel strategy is `Threads.@threads for k` (arrowhead3.jl:641): eigenpairs are
independent given (D,z,a), so every rank gets the (tiny) inputs and owns a contiguous block of k and
of U's columns.  No collective is needed to compute; `allgather_columns` (NCCL over NVLink via
torch.distributed) exists for callers that want all of U on every device.
"""
from __future__ import annotations

import numpy as np


def k_partition(nk: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous 1-based inclusive range [k0,k1] of `nk` eigenpairs owned by `rank`; ranges differ by
    at most one eigenpair; empty ranges (nk < world) come back as (1, 0)."""
    base, rem = divmod(int(nk), int(world))
    cnt = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    if cnt == 0:
        return 1, 0
    return start + 1, start + cnt


def solve_sharded(ctx, kind: str, args: tuple, nk: int, rank: int, world: int, **kw):
    """Solve this rank's k-range.  kind: 'symarrow' | 'symdpr1' | 'halfarrow'; args as for Context.<kind>."""
    k0, k1 = k_partition(nk, world, rank)
    if k1 < k0:
        return None
    return getattr(ctx, kind)(*args, k0=k0, k1=k1, **kw)


def gather_values(local: np.ndarray | None, nk: int, world: int, rank: int, dist=None) -> np.ndarray:
    """All-gather the per-k host arrays (eigenvalues / Info fields) over the process group."""
    if dist is None or world == 1:
        return np.asarray(local)
    import torch
    counts = [k_partition(nk, world, r) for r in range(world)]
    maxc = max(k1 - k0 + 1 for k0, k1 in counts)
    buf = torch.zeros(maxc, dtype=torch.float64)
    if local is not None and len(local):
        buf[: len(local)] = torch.from_numpy(np.asarray(local, np.float64))
    outs = [torch.zeros(maxc, dtype=torch.float64) for _ in range(world)]
    if dist.get_backend() == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        buf = buf.to(dev)
        outs = [o.to(dev) for o in outs]
    dist.all_gather(outs, buf)
    return np.concatenate([o.cpu().numpy()[: (k1 - k0 + 1)] for o, (k0, k1) in zip(outs, counts)])


def allgather_columns(U_local, nrows: int, nk: int, world: int, rank: int, dist):
    """All-gather column blocks of a device-resident U (torch CUDA tensor of shape [cnt, nrows], i.e. the
    column-major n x cnt block) into one [nk, nrows] tensor on every rank (NCCL all_gather over NVLink)."""
    import torch
    counts = [k_partition(nk, world, r) for r in range(world)]
    maxc = max(k1 - k0 + 1 for k0, k1 in counts)
    pad = torch.zeros((maxc, nrows), dtype=torch.float64, device=U_local.device)
    pad[: U_local.shape[0]] = U_local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[: (k1 - k0 + 1)] for o, (k0, k1) in zip(outs, counts)], dim=0)
