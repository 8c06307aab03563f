"""world_size-2 check of the sharding plumbing on the CPU (gloo): k-ranges tile 1..n and the per-k host
arrays are gathered in order.  The per-rank solver is replaced by a stand-in (f(k) = k^2) because the
real one needs a GPU; the multi-GPU numerics are covered by the gpu tests and bench.py --gpus N."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nk, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import arrowhead_b200 as ab
    from arrowhead_b200.sharding import gather_values
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    k0, k1 = ab.k_partition(nk, world, rank)
    local = np.array([float(k) ** 2 for k in range(k0, k1 + 1)])
    full = gather_values(local, nk, world, rank, dist)
    # column blocks of U (row c of the [cnt, nrows] tensor = eigenvector column k0+c), ragged split
    import torch
    from arrowhead_b200.sharding import allgather_columns
    nrows = 5
    Uloc = torch.tensor([[100.0 * k + r for r in range(nrows)] for k in range(k0, k1 + 1)], dtype=torch.float64)
    Ufull = allgather_columns(Uloc, nrows, nk, world, rank, dist)
    assert Ufull.shape == (nk, nrows)
    assert torch.equal(Ufull, torch.tensor([[100.0 * k + r for r in range(nrows)] for k in range(1, nk + 1)], dtype=torch.float64))
    q.put((rank, k0, k1, full.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gather_world2():
    world, nk = 2, 11
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nk, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [float(k) ** 2 for k in range(1, nk + 1)]
    ranges = sorted((k0, k1) for _, k0, k1, _ in res)
    assert ranges == [(1, 6), (7, 11)]
    for _, _, _, full in res:
        assert full == expect
