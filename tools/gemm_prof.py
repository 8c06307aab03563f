#!/usr/bin/env python
"""Workload for ncu: one ah_dgemm_batched at the top back-transform shape of tdc n = 8192 and the cuBLAS call beside it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import arrowhead_b200 as ab
ctx = ab.get_context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
m, n, k = 4096, 8192, 4096
A = torch.randn((k, m), dtype=torch.float64, device="cuda"); B = torch.randn((n, k), dtype=torch.float64, device="cuda")
C = torch.empty((n, m), dtype=torch.float64, device="cuda")
for _ in range(2):
    ctx.dgemm_batched(m, n, k, [A.data_ptr()], m, [B.data_ptr()], k, [C.data_ptr()], m)
    torch.mm(B, A, out=C)
torch.cuda.synchronize()
