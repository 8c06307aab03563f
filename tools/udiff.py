#!/usr/bin/env python
"""Bitwise comparison of the vectors two builds of the library produce: python tools/udiff.py libA.so libB.so"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
import arrowhead_b200 as ab
from helpers import gen_symarrow, gen_halfarrow, gen_symdpr1
ctx = ab.get_context(0)
D, z, a = gen_symarrow(4096, seed=7)
r = ctx.symarrow(D, z, a)
Dh, zh = gen_halfarrow(2047, True, seed=7)
h = ctx.halfarrow(Dh, zh)
Dd, ud, rho = gen_symdpr1(3000, seed=7)
d = ctx.symdpr1(Dd, ud, rho)
np.savez(sys.argv[1], U=r.U, HU=h.U, HV=h.V, DU=d.U)
''' % (ROOT, ROOT)
for i, lib in enumerate(sys.argv[1:3]):
    subprocess.run([sys.executable, "-c", code, f"/tmp/U{i}.npz"], env=dict(os.environ, ARROWHEAD_B200_LIB=os.path.abspath(lib)), check=True)
import numpy as np
A, B = np.load("/tmp/U0.npz"), np.load("/tmp/U1.npz")
for key in A.files:
    a, b = A[key], B[key]
    d = a != b
    rel = np.abs(a - b)[d] / np.abs(a)[d] if d.any() else np.zeros(1)
    print(key, "entries", a.size, "differing", int(d.sum()), "max rel diff / eps", float(rel.max() / 2.220446049250313e-16))
