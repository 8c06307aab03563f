#!/usr/bin/env python
"""Kernel-time breakdown with the accelerated bisection on / off (ARROWHEAD_B200_ACCEL), full launches and a 1/8 k-range."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import arrowhead_b200 as ab
import bench

cases = [("symarrow", 32768, 1), ("symarrow", 32768, 8), ("symarrow", 65536, 1), ("symdpr1", 16384, 1), ("halfarrow", 16384, 1), ("symarrow", 8192, 1)]
if len(sys.argv) > 1:
    cases = [(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))]
for kind, n, world in cases:
    P = bench.Problem(kind, n)
    k0, k1 = ab.k_partition(n, world, 0)
    U = torch.empty((k1 - k0 + 1, n), dtype=torch.float64, device="cuda")
    V = torch.empty((k1 - k0 + 1, n), dtype=torch.float64, device="cuda") if kind == "halfarrow" else None
    for accel in ("0", "1"):
        os.environ["ARROWHEAD_B200_ACCEL"] = accel
        ctx = ab.Context(0)
        best = None
        for rep in range(4):
            r = P.solve(ctx, k0, k1, U.data_ptr(), None if V is None else V.data_ptr())
            s = r.stats
            if rep and (best is None or s["kernel_ms"] < best["kernel_ms"]):
                best = dict(s)
        cnt = k1 - k0 + 1
        print(f"{kind:9s} n={n}/{world} accel={accel}: kernel={best['kernel_ms']:.3f} bisect={best['bisect_ms']:.3f} prep={best['prep_ms']:.3f} "
              f"vec={best['vec_ms']:.3f} rem3={best['rem3_ms']:.3f} full={best['full_ms']:.3f} sweeps/eigenpair={r.steps.sum() / cnt:.1f} "
              f"n_rem3={best['n_rem3']} n_full={best['n_full']} -> {cnt / best['kernel_ms'] / 1e3:.3f} M eigenpairs/s", flush=True)
        ctx.close()
        del ctx
