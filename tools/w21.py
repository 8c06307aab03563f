#!/usr/bin/env python
"""W21 (runtests.jl:91-115): ||sort(L) - sort(L_known)||_2 in eps for the three tdc paths."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import arrowhead_b200 as ab, oracle
EPS = 2.0 ** -52
raw = json.load(open(os.path.join(ROOT, "tests", "golden", "runtests_known_answers.json")))["tdc_w21"]
hx = lambda v: np.array([float.fromhex(x) for x in v])
dv, ev, vals = hx(raw["dv"]), hx(raw["ev"]), hx(raw["values"])
ctx = ab.get_context(0)
for name, fn in (("tdc_recursive (host back-transform)", lambda: ab.tdc_recursive(dv, ev, ctx=ctx).values),
                 ("tdc_device (DMMA back-transform)", lambda: ab.tdc_device(dv, ev, ctx=ctx).values),
                 ("oracle.tdc (CPU restatement)", lambda: oracle.tdc(dv, ev)[0])):
    L = fn()
    d = np.sort(L) - np.sort(vals)
    print(f"{name:40s} ||diff||_2 = {np.linalg.norm(d) / EPS:7.2f} eps   max |diff|/|lambda| = {np.max(np.abs(d) / np.abs(np.sort(vals))) / EPS:5.2f} eps")
