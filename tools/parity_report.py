#!/usr/bin/env python
"""Parity numbers for DESIGN.md: GPU vs oracle at the BASELINE sizes (all k; vectors on a column sample)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from types import SimpleNamespace
import arrowhead_b200 as ab, oracle
from helpers import EPS, compare_range, gen_symarrow, gen_symdpr1, shift_margin_symarrow, shift_margin_symdpr1
ctx = ab.get_context(0); olib = oracle.get_lib(); ohi = oracle.get_lib_hi()
out = {}
for n in (1000, 32768):
    D, z, a = gen_symarrow(n, seed=421)
    U = torch.empty((n, n), dtype=torch.float64, device="cuda")
    G = ctx.symarrow(D, z, a, vectors=False, U_dev=U.data_ptr(), ldu=n)
    tau = oracle.default_tau(n - 1)
    O = olib.symarrow(D, z, a, tau, vectors=False); H = ohi.symarrow(D, z, a, tau, vectors=False)
    r = compare_range(SimpleNamespace(k0=1, lam=G.lam, sind=G.sind, qout=G.qout, status=G.status, kb=G.kb, kz=G.kz, knu=G.knu,
                                      krho=G.krho, U=None, V=None), O, H, k_tol=1e-7, margin_fn=lambda k: shift_margin_symarrow(D, z, a, k))
    lam_err = np.abs(G.lam - O.lam) / np.abs(O.lam) / EPS
    r.update(lam_err_p50=float(np.median(lam_err)), lam_err_p999=float(np.percentile(lam_err, 99.9)), n_rem3=int((O.krho != 0).sum()),
             sind_equal=int((G.sind == O.sind).sum()), qout_equal=int((G.qout == O.qout).sum()))
    rng = np.random.default_rng(0)
    ks = np.unique(rng.integers(1, n + 1, 200))
    within = []; worst = 0.0; worst_hi = 0.0
    for k in ks:
        k = int(k)
        Ok = olib.symarrow(D, z, a, tau, k0=k, k1=k); Hk = ohi.symarrow(D, z, a, tau, k0=k, k1=k)
        g = U[k - 1].cpu().numpy()
        e = np.max(np.abs(g - Ok.U[:, 0]) / np.abs(Ok.U[:, 0])) / EPS
        eh = np.max(np.abs(g - Hk.U[:, 0]) / np.abs(Hk.U[:, 0])) / EPS
        oh = np.max(np.abs(Ok.U[:, 0] - Hk.U[:, 0]) / np.abs(Hk.U[:, 0])) / EPS
        within.append((e, eh, oh))
    w = np.array(within)
    r.update(vec_columns_sampled=int(ks.size), vec_frac_within_100eps=float((w[:, 0] <= 100).mean()), vec_err_gpu_vs_oracle_p50=float(np.median(w[:, 0])),
             vec_err_gpu_vs_oracle_max=float(w[:, 0].max()), vec_err_gpu_vs_exact_sums_p50=float(np.median(w[:, 1])), vec_err_gpu_vs_exact_sums_max=float(w[:, 1].max()),
             vec_err_oracle_vs_exact_sums_p50=float(np.median(w[:, 2])), vec_err_oracle_vs_exact_sums_max=float(w[:, 2].max()))
    out[f"symarrow_n{n}"] = r
    del U
print(json.dumps(out, indent=1))
