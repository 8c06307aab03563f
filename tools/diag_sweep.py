#!/usr/bin/env python
"""Cycles per sweep of k_bisect's warp 0 (AH_DIAG_TIMING build in variants/lib_diag.so), accelerated vs plain bisection."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import arrowhead_b200 as ab, bench
lib = ab.load_library(os.path.join(ROOT, "variants", "lib_diag.so"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
P = bench.Problem("symarrow", n)
U = torch.empty((n, n), dtype=torch.float64, device="cuda")
for accel in ("0", "1"):
    os.environ["ARROWHEAD_B200_ACCEL"] = accel
    ctx = ab.Context(0, lib=lib)
    print(f"== ACCEL={accel}", flush=True)
    for _ in range(2):
        r = P.solve(ctx, 1, n, U.data_ptr())
    print(r.stats["bisect_ms"], r.steps.mean(), flush=True)
    ctx.close()
