"""Every non-standard branch of the per-k solver is exercised, and the CUDA path takes the SAME branch as the oracle.

tests/golden/branch_cases.json (made by tests/golden/make_branch_cases.py) holds small inputs on which the oracle goes
through: double-double b, Remedy 3, repeated Remedy 3, the 'R'->'L' retry, double-double rho, Remedy 1, Remedy 1 with
rho = Inf, |nu| = Inf -- for each matrix type where the branch exists.  Both the oracle and the library record the
branch in bits 8..14 of the per-eigenpair status word (AH_BR_* in include/arrowhead_b200.h).

CPU: the stored inputs still fire the stored branches in the oracle (pins fixture <-> oracle).
GPU: status words (conditions + branch trace), shift index and Qout equal the oracle's for every eigenpair of every
case, values agree, and every (type, branch) target is hit at least once in k_full."""
import json
import os

import numpy as np
import pytest

import oracle
from helpers import EPS

HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    with open(os.path.join(HERE, "golden", "branch_cases.json"), encoding="utf-8") as f:
        raw = json.load(f)
    cases = []
    for c in raw["cases"]:
        D = np.array([float.fromhex(x) for x in c["D"]]); w = np.array([float.fromhex(x) for x in c["w"]])
        top = None if c["top"] is None else float.fromhex(c["top"])
        cases.append((c["kind"], c["fires"], D, w, top, np.array(c["status"], np.int32), np.array(c["qout"], np.int64)))
    return raw["bits"], raw["not_found"], cases


BITS, NOT_FOUND, CASES = _load()
EXPECTED = {
    "symarrow": {"DD_B", "REM3", "REM3_REPEAT", "RL_RETRY", "DD_RHO", "REM1", "REM1_ZERO", "NU_INF"},
    "symdpr1": {"DD_B", "REM3", "REM3_REPEAT", "DD_RHO", "REM1", "NU_INF"},
    "halfarrow": {"DD_B", "REM3", "DD_RHO", "REM1", "NU_INF"},
}


def _oracle(lib, kind, D, w, top, vectors=False):
    tau = oracle.default_tau(D.size)
    if kind == "symarrow":
        return lib.symarrow(D, w, top, tau, vectors=vectors)
    if kind == "symdpr1":
        return lib.symdpr1(D, w, top, tau, vectors=vectors)
    return lib.halfarrow(D, w, tau, vectors=vectors)


def test_fixture_covers_every_branch_of_every_type():
    assert NOT_FOUND == []
    seen = {k: set() for k in EXPECTED}
    for kind, fires, *_ in CASES:
        seen[kind] |= set(fires)
    assert seen == EXPECTED


def test_oracle_takes_the_stored_branches(oracle_lib):
    for kind, fires, D, w, top, status, qout in CASES:
        O = _oracle(oracle_lib, kind, D, w, top)
        assert np.array_equal(O.status, status) and np.array_equal(O.qout, qout), (kind, fires)
        for f in fires:
            assert ((O.status & BITS[f]) != 0).any(), (kind, f)


@pytest.mark.gpu
def test_gpu_takes_the_same_branches_as_the_oracle(gpu_ctx, oracle_lib):
    hits = {k: {b: 0 for b in EXPECTED[k]} for k in EXPECTED}
    for kind, fires, D, w, top, status, qout in CASES:
        O = _oracle(oracle_lib, kind, D, w, top, vectors=True)
        if kind == "symarrow":
            G = gpu_ctx.symarrow(D, w, top)
        elif kind == "symdpr1":
            G = gpu_ctx.symdpr1(D, w, top)
        else:
            G = gpu_ctx.halfarrow(D, w)
        tag = (kind, fires, D.size)
        assert np.array_equal(G.sind, O.sind), tag
        assert np.array_equal(G.status & ~32, O.status), (tag, [hex(s) for s in G.status], [hex(s) for s in O.status])
        assert np.array_equal(G.qout, O.qout), tag
        for b in EXPECTED[kind]:
            hits[kind][b] += int(((G.status & BITS[b]) != 0).sum())
        # values: where the reference's result is well defined (no |nu| = Inf / BigFloat / pole-return / raise flags)
        fine = ((O.status & 0xff) == 0) & np.isfinite(O.lam) & ((G.status & 32) == 0)
        den = np.where(O.lam == 0, 1.0, np.abs(O.lam))
        err = np.abs(G.lam - O.lam) / den / EPS
        # after a Remedy the result carries the K_nu / K_rho-fold noise of the first stage (tests/helpers.py): the bar
        # is 10 eps on plain eigenpairs and widened by the oracle's own condition numbers elsewhere
        knu = np.where(np.isfinite(O.knu), O.knu, 0.0); krho = np.where(np.isfinite(O.krho), np.minimum(O.krho, 1 / EPS), 1 / EPS)
        bar = 10 + 0.1 * knu + np.where(O.krho != 0, krho, 0.0)
        assert np.all(err[fine] <= bar[fine]), (tag, err[fine], bar[fine])
        nu_inf = (O.status & 1) != 0                                   # deflation branch: lambda = sigma exactly, v = e_i
        assert np.array_equal(G.lam[nu_inf], O.lam[nu_inf]), tag
        if kind != "halfarrow" and nu_inf.any():
            assert np.array_equal(G.U[:, nu_inf], O.U[:, nu_inf]), tag
    for kind in hits:
        for b, cnt in hits[kind].items():
            assert cnt > 0, f"{kind}: branch {b} never fired on the GPU"
