import sys, os, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import arrowhead_b200 as ab
from helpers import gen_symarrow
ctx = ab.get_context(0)
n = 32768
D, z, a = gen_symarrow(n, seed=421)
for world in (1, 2, 4, 8):
    k0, k1 = ab.k_partition(n, world, 0)
    cnt = k1 - k0 + 1
    U = torch.empty((cnt, n), dtype=torch.float64, device="cuda")
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        G = ctx.symarrow(D, z, a, k0=k0, k1=k1, vectors=False, U_dev=U.data_ptr(), ldu=n)
        wall = (time.perf_counter() - t0) * 1e3
    s = G.stats
    print(f"world={world} cnt={cnt} wall={wall:.2f} ms kernel={s['kernel_ms']:.2f} ms bisect={s['bisect_ms']:.2f} prep={s['prep_ms']:.2f} vec={s['vec_ms']:.2f} -> scaled eff vs world=1: ", end="")
    if world == 1: base = s['kernel_ms']
    print(f"{base / (world * s['kernel_ms']):.3f}")
    del U
