#!/usr/bin/env python
"""Where does a bench step's time go besides the kernels?  python tools/step_overhead.py [n] [world]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import arrowhead_b200 as ab
from helpers import gen_symarrow

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = ab.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
D, z, a = gen_symarrow(n)
k0, k1 = ab.k_partition(n, world, 0)
U = torch.empty((k1 - k0 + 1, n), dtype=torch.float64, device="cuda")
flush = torch.empty(32 << 20, dtype=torch.float64, device="cuda")     # 256 MB
fe0, fe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush.zero_(); fe0.record(); flush.zero_(); fe1.record(); torch.cuda.synchronize()
print("flush.zero_() of 256 MB as float64: %.3f ms" % fe0.elapsed_time(fe1))
for _ in range(3):
    ctx.symarrow(D, z, a, k0=k0, k1=k1, vectors=False, U_dev=U.data_ptr(), ldu=n)
torch.cuda.synchronize()
K = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for with_flush in (True, False):
    kern = 0.0; wall_call = 0.0
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record()
    for _ in range(K):
        if with_flush:
            flush.zero_()
        c0 = time.perf_counter()
        r = ctx.symarrow(D, z, a, k0=k0, k1=k1, vectors=False, U_dev=U.data_ptr(), ldu=n)
        wall_call += time.perf_counter() - c0
        kern += r.stats["kernel_ms"]
    e1.record(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"n={n}/{world} flush={with_flush}: events {e0.elapsed_time(e1)/K:.3f} ms/step, wall {(t1-t0)/K*1e3:.3f}, "
          f"symarrow() wall {wall_call/K*1e3:.3f}, kernels {kern/K:.3f}, stats={ {k: round(v,3) for k,v in r.stats.items() if k.endswith('_ms')} }")

# ---- the same steps while NVML is polled from a second thread (bench.py's clock sampler)
if len(sys.argv) > 3:
    import threading
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    for interval in (0.01, 0.05):
        smp = bench.ClockSampler(0)
        smp.interval = interval
        smp.start()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(K * 4):
            flush.zero_()
            r = ctx.symarrow(D, z, a, k0=k0, k1=k1, vectors=False, U_dev=U.data_ptr(), ldu=n)
        e1.record(); torch.cuda.synchronize()
        smp.stop_flag = True; smp.join()
        print(f"with NVML sampler every {interval * 1e3:.0f} ms: events {e0.elapsed_time(e1) / (K * 4):.3f} ms/step ({len(smp.samples)} samples)")
