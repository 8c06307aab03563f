#!/usr/bin/env python
"""profiles/r2_parity_report.json: GPU-vs-truth and oracle-vs-truth side by side for every mpmath fixture
(tests/golden/truth_*.npz), plus the all-k integer-field comparison with the oracle at the BASELINE sizes.
Run on the GPU box:  python tools/parity_report.py > gpurun_out/r2_parity_report.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import arrowhead_b200 as ab
import oracle
from test_gpu_truth import measure_case, summarize
from truth import Truth, available

ctx = ab.get_context(0)
olib = oracle.get_lib()
out = {"truth": "320-bit mpmath roots of the secular equations, certified by sign change (tests/golden/make_truth.py)",
       "units": "eps = 2^-52, relative; eigenvector error = max over the entries of a column"}
for name in ("symarrow_n1000", "symarrow_n32768", "symdpr1_n16384", "halfarrow_n16384", "symarrow_n65536"):
    if not available(name):
        continue
    T = Truth(name)
    m, G = measure_case(ctx, T, olib)
    s = summarize(m)
    # all-k integer fields against the oracle (vectors off: seconds on the host cores)
    n = T.n
    if T.kind == "symarrow":
        O = olib.symarrow(T.D, T.z, T.a, oracle.default_tau(n - 1), vectors=False)
    elif T.kind == "symdpr1":
        O = olib.symdpr1(T.D, T.u, T.rho, oracle.default_tau(n), vectors=False)
    else:
        O = olib.halfarrow(T.D, T.z, oracle.default_tau(n - 1), vectors=False)
    s["all_k"] = {"n": int(G.lam.size), "shift_index_equal": int((G.sind == O.sind).sum()), "qout_equal": int((G.qout == O.qout).sum()),
                  "remedy3_same_set": bool(np.array_equal(G.krho != 0, O.krho != 0)), "n_remedy3": int((O.krho != 0).sum()),
                  "status_nonzero": int((G.status != 0).sum()),
                  "lam_gpu_vs_oracle_eps_max": float(np.max(np.abs(G.lam - O.lam) / np.abs(O.lam)) / 2.0 ** -52)}
    worst = int(np.argmax(m["vec_gpu"]))
    s["worst_gpu_column"] = {"k": int(m["k"][worst]), "gpu": float(m["vec_gpu"][worst]), "oracle": float(m["vec_orc"][worst]),
                             "knu": float(m["knu"][worst])}
    out[name] = s
print(json.dumps(out, indent=1))
