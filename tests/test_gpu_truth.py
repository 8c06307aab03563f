"""GPU path vs the independent mpmath truth (tests/golden/truth_*.npz) -- VERDICT r1 "Next round" item 1.

For every eigenpair of every fixture (all 1000 of configs[0]; 256 sampled k of configs[1..3], 64 of the n = 65536 target):

    err(GPU, truth)  <=  max( err(oracle, truth),  10 eps [eigenvalue] | 100 eps [eigenvector entries] )

i.e. the CUDA path is never further from the TRUE eigenpair than the reference's own algorithm is, and wherever the
reference meets the north-star bars the CUDA path meets them too.  On configs[0] the plain bars are asserted for 100 %
of the eigenpairs outright.  The integer Info fields (shift index, Qout) must equal the oracle's for every sampled k."""
import numpy as np
import pytest

import oracle
from helpers import EPS
from truth import Truth, available, err_eps

pytestmark = pytest.mark.gpu


def measure_case(ctx, T, olib, cols=None):
    """Solve the whole problem on the GPU (vectors resident), then compare the fixture's eigenpairs with the truth and
    with the oracle's single-k solves.  Returns per-eigenpair error arrays (eps units)."""
    import torch
    n = T.n
    cols = list(range(len(T.k))) if cols is None else list(cols)
    if T.kind == "symarrow":
        U = torch.empty((n, n), dtype=torch.float64, device="cuda")
        G = ctx.symarrow(T.D, T.z, T.a, vectors=False, U_dev=U.data_ptr(), ldu=n)
        V = None
    elif T.kind == "symdpr1":
        U = torch.empty((n, n), dtype=torch.float64, device="cuda")
        G = ctx.symdpr1(T.D, T.u, T.rho, vectors=False, U_dev=U.data_ptr(), ldu=n)
        V = None
    else:
        U = torch.empty((n, n), dtype=torch.float64, device="cuda")
        V = torch.empty((n, n), dtype=torch.float64, device="cuda")
        G = ctx.halfarrow(T.D, T.z, vectors=False, U_dev=U.data_ptr(), ldu=n, V_dev=V.data_ptr(), ldv=n)
    torch.cuda.synchronize()
    out = {key: [] for key in ("k", "lam_gpu", "lam_orc", "vec_gpu", "vec_orc", "left_gpu", "left_orc", "knu", "krho",
                               "sind_equal", "qout_equal", "status")}
    for c in cols:
        k = int(T.k[c])
        if T.kind == "symarrow":
            O = olib.symarrow(T.D, T.z, T.a, oracle.default_tau(n - 1), k0=k, k1=k)
            gvec, ovec = U[k - 1].cpu().numpy(), O.U[:, 0]
        elif T.kind == "symdpr1":
            O = olib.symdpr1(T.D, T.u, T.rho, oracle.default_tau(n), k0=k, k1=k)
            gvec, ovec = U[k - 1].cpu().numpy(), O.U[:, 0]
        else:
            O = olib.halfarrow(T.D, T.z, oracle.default_tau(n - 1), k0=k, k1=k)
            gvec, ovec = V[k - 1].cpu().numpy(), O.V[:, 0]
            tl = T.left_vector(c)
            out["left_gpu"].append(float(err_eps(U[k - 1].cpu().numpy(), tl).max()))
            out["left_orc"].append(float(err_eps(O.U[:, 0], tl).max()))
        tv = T.vector(c)
        out["k"].append(k)
        out["lam_gpu"].append(float(err_eps(G.lam[k - 1], T.lam[c])))
        out["lam_orc"].append(float(err_eps(O.lam[0], T.lam[c])))
        out["vec_gpu"].append(float(err_eps(gvec, tv).max()))
        out["vec_orc"].append(float(err_eps(ovec, tv).max()))
        out["knu"].append(float(O.knu[0])); out["krho"].append(float(O.krho[0]))
        out["sind_equal"].append(bool(G.sind[k - 1] == O.sind[0]))
        out["qout_equal"].append(bool(G.qout[k - 1] == O.qout[0]))
        out["status"].append(int(G.status[k - 1]))
    del U, V
    return {key: np.array(v) for key, v in out.items()}, G


def summarize(m):
    s = {"eigenpairs": int(m["k"].size)}
    for side in ("gpu", "orc"):
        for q in ("lam", "vec") + (("left",) if m["left_gpu"].size else ()):
            e = m[f"{q}_{side}"]
            bar = 10.0 if q == "lam" else 100.0
            s[f"{q}_{side}_vs_truth_eps"] = {"median": float(np.median(e)), "p99": float(np.percentile(e, 99)), "max": float(e.max()),
                                            "frac_within_bar": float((e <= bar).mean())}
    s["gpu_never_worse_than_max_oracle_bar"] = {
        "lam": bool(np.all(m["lam_gpu"] <= np.maximum(m["lam_orc"], 10.0))),
        "vec": bool(np.all(m["vec_gpu"] <= np.maximum(m["vec_orc"], 100.0)))}
    s["shift_index_equal"] = int(m["sind_equal"].sum())
    s["qout_equal"] = int(m["qout_equal"].sum())
    s["remedy3_in_sample"] = int((m["krho"] != 0).sum())
    s["max_knu_in_sample"] = float(m["knu"].max())
    return s


def _assert_case(m, plain_100pct=False):
    assert m["sind_equal"].all(), f"shift index differs from the oracle at k={m['k'][~m['sind_equal']]}"
    assert m["qout_equal"].all(), f"Qout differs from the oracle at k={m['k'][~m['qout_equal']]}"
    assert ((m["status"] & 0xff) == 0).all()
    bad = m["lam_gpu"] > np.maximum(m["lam_orc"], 10.0)
    assert not bad.any(), f"eigenvalue further from the truth than the reference's at k={m['k'][bad]}: " \
                          f"{m['lam_gpu'][bad]} vs {m['lam_orc'][bad]} eps"
    bad = m["vec_gpu"] > np.maximum(m["vec_orc"], 100.0)
    assert not bad.any(), f"eigenvector further from the truth than the reference's at k={m['k'][bad]}: " \
                          f"{m['vec_gpu'][bad]} vs {m['vec_orc'][bad]} eps"
    if m["left_gpu"].size:
        bad = m["left_gpu"] > np.maximum(m["left_orc"], 100.0)
        assert not bad.any(), f"left singular vector further from the truth than the reference's at k={m['k'][bad]}: " \
                              f"{m['left_gpu'][bad]} vs {m['left_orc'][bad]} eps"
    if plain_100pct:
        assert m["lam_gpu"].max() <= 10.0 and m["vec_gpu"].max() <= 100.0, (m["lam_gpu"].max(), m["vec_gpu"].max())


def test_config0_plain_bars_for_every_eigenpair(gpu_ctx, oracle_lib):
    """BASELINE configs[0] (GenSymArrow(1000,1000)-sized): 10 eps / 100 eps against the TRUTH for all 1000 eigenpairs --
    the reference itself misses the 100-eps bar on 2 columns here (tests/test_truth.py)."""
    T = Truth("symarrow_n1000")
    m, G = measure_case(gpu_ctx, T, oracle_lib)
    _assert_case(m, plain_100pct=True)
    assert m["k"].size == 1000


@pytest.mark.parametrize("name", ["symarrow_n32768", "symdpr1_n16384", "halfarrow_n16384", "symarrow_n65536"])
def test_gpu_is_never_further_from_the_truth_than_the_reference(gpu_ctx, oracle_lib, name):
    assert available(name)
    T = Truth(name)
    m, G = measure_case(gpu_ctx, T, oracle_lib)
    _assert_case(m)
    # and in absolute terms.  SymArrow / SymDPR1: the plain north-star bars hold against the truth for EVERY sampled
    # eigenpair (the reference itself is up to 2750 eps off there).  HalfArrow: the three smallest singular values of
    # the sample (sigma ~ 1e-5 ||A||, first-stage K_nu ~ 1e9) are beyond what the reference's squared formulation can
    # deliver -- the reference is 4e3..4e5 eps off, the CUDA path 4e2..2e4 -- so there the bar is "at least as often
    # within the bars as the reference, and >= 98 %".
    if T.kind == "halfarrow":
        assert (m["lam_gpu"] <= 10.0).mean() >= max(0.98, (m["lam_orc"] <= 10.0).mean()), summarize(m)
        assert (m["vec_gpu"] <= 100.0).mean() >= max(0.98, (m["vec_orc"] <= 100.0).mean()), summarize(m)
    else:
        assert m["lam_gpu"].max() <= 10.0 and m["vec_gpu"].max() <= 100.0, summarize(m)
