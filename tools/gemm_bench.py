#!/usr/bin/env python
"""ah_dgemm_batched (own DMMA kernel) vs cuBLAS (torch) on the back-transform shapes of tdc n=8192: python tools/gemm_bench.py [lib.so]"""
import os, sys
if len(sys.argv) > 1:
    os.environ["ARROWHEAD_B200_LIB"] = os.path.abspath(sys.argv[1])
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import arrowhead_b200 as ab
ctx = ab.get_context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
print(os.path.basename(os.environ.get("ARROWHEAD_B200_LIB", "libarrowhead_b200.so")), "DMMA burst %.1f TFLOP/s" % (2 * ctx.measure_dmma_peak()[0] / 1e3))
for nb, m, n, k in ((1, 4096, 8192, 4096), (1, 4095, 8192, 4095), (1, 2048, 4096, 2048), (3, 1023, 2047, 1023), (7, 511, 1023, 511),
                    (15, 255, 511, 255), (31, 127, 255, 127), (63, 63, 127, 63), (255, 15, 31, 15)):
    A = torch.randn((nb, k, m), dtype=torch.float64, device="cuda"); B = torch.randn((nb, n, k), dtype=torch.float64, device="cuda")
    C = torch.empty((nb, n, m), dtype=torch.float64, device="cuda")
    pa = [A[b].data_ptr() for b in range(nb)]; pb = [B[b].data_ptr() for b in range(nb)]; pc = [C[b].data_ptr() for b in range(nb)]
    own = lambda: ctx.dgemm_batched(m, n, k, pa, m, pb, k, pc, m)
    blas = lambda: torch.bmm(B, A, out=C)
    res = []
    for fn in (own, blas):
        fn(); torch.cuda.synchronize(); best = 1e30
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        res.append(best)
    fl = 2.0 * nb * m * n * k
    print(f"  nb={nb:4d} m={m:5d} n={n:5d} k={k:5d}: own {res[0]:8.3f} ms {fl / res[0] / 1e9:6.2f} TF | cuBLAS {res[1]:8.3f} ms {fl / res[1] / 1e9:6.2f} TF | ratio {res[1] / res[0]:.3f}", flush=True)
