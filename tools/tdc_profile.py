#!/usr/bin/env python
"""Where tdc_device spends its time: every C-ABI call timed synchronously (diagnostic; the production path is async)."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import arrowhead_b200 as ab
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rng = np.random.default_rng(421)
dv, ev = 2.0 + 1e-3 * rng.standard_normal(n), -1.0 + 1e-3 * rng.standard_normal(n - 1)
ctx = ab.get_context(0)
E = ab.tdc_device(dv, ev, ctx=ctx, vectors="device"); E.vectors.free()
tot = collections.defaultdict(float); cnt = collections.Counter()
def wrap(name):
    fn = getattr(ctx, name)
    def w(*a, **k):
        t0 = time.perf_counter(); r = fn(*a, **k); ctx.wait(); dt = time.perf_counter() - t0
        tot[name] += dt; cnt[name] += 1
        if name in ("symarrow_batched", "symarrow"):
            key = f"{name}[N={np.asarray(a[0]).shape[-1]}]"; tot[key] += dt; cnt[key] += 1
        return r
    setattr(ctx, name, w)
for name in ("symarrow", "symarrow_batched", "gather_elements", "gather_permuted_batched", "copy2d_batched", "dgemm_batched", "set_identity", "gather_permuted"):
    wrap(name)
ctx.set_async = lambda on: None
t0 = time.perf_counter()
E = ab.tdc_device(dv, ev, ctx=ctx, vectors="device"); ctx.wait()
total = time.perf_counter() - t0
E.vectors.free()
print(f"tdc n={n}: {total * 1e3:.1f} ms with every call synchronised")
acc = 0.0
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"  {k:36s} {cnt[k]:5d} calls {v * 1e3:8.2f} ms")
    if "[" not in k: acc += v
print(f"  {'host logic between the calls':36s}       {(total - acc) * 1e3:8.2f} ms")
