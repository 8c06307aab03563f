import sys, time
sys.path.insert(0, ".")
import numpy as np, arrowhead_b200 as ab
ctx = ab.get_context(0)
n = 8192
rng = np.random.default_rng(421)
dv, ev = 2.0 + 1e-3 * rng.standard_normal(n), -1.0 + 1e-3 * rng.standard_normal(n - 1)
for rep in range(3):
    t0 = time.time(); E = ab.tdc_device(dv, ev, ctx=ctx, vectors="device"); dt = time.time() - t0; E.vectors.free()
    print("tdc_device", round(dt * 1e3, 1), "ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); E = ab.tdc_device(dv, ev, ctx=ctx, vectors="device"); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
