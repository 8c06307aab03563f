"""The accelerated bisection of k_bisect: grid points of the reference's bisection are decided from
certified bounds instead of a sweep.  The certificates rest on (a) an error bound of the reciprocal seed, measured
exhaustively here, and (b) monotonicity arguments; the proof of the pudding is that every output bit equals plain
bisection's (ARROWHEAD_B200_ACCEL=0) on every family of the parity and stress tests."""
import numpy as np
import pytest

import arrowhead_b200 as ab
from helpers import gen_halfarrow, gen_symarrow, gen_symdpr1
from test_gpu_stress import families as arrow_families
from test_gpu_stress_half import families as half_families

pytestmark = pytest.mark.gpu


def test_reciprocal_seed_error_is_inside_the_certificates_assumption(gpu_ctx):
    seed, corrected = gpu_ctx.measure_rcp_seed_error()
    print(f"MUFU.RCP64H: max |1 - x r0| = 2^{np.log2(seed):.2f}; corrected reciprocal: {corrected / 2.0 ** -53:.3f} x 2^-53")
    assert 0 < seed <= 2.0 ** -18          # theta of the certificates assumes eps^3 <= 2^-54
    assert 0 < corrected <= 1.6 * 2.0 ** -53


def _contexts(monkeypatch):
    out = []
    for accel in ("0", "1"):
        monkeypatch.setenv("ARROWHEAD_B200_ACCEL", accel)
        out.append(ab.Context(0))            # the knob is read when the context is created
    monkeypatch.delenv("ARROWHEAD_B200_ACCEL", raising=False)
    return out


FIELDS = ("lam", "sind", "kb", "kz", "knu", "krho", "qout", "status")


def _same(r0, r1, what):
    for f in FIELDS:
        assert np.array_equal(getattr(r0, f), getattr(r1, f), equal_nan=True), (what, f)
    for f in ("U", "V"):
        a, b = getattr(r0, f, None), getattr(r1, f, None)
        if a is not None:
            assert np.array_equal(a, b, equal_nan=True), (what, f)


@pytest.mark.timeout(600)
def test_accelerated_bisection_is_bitwise_plain_bisection(monkeypatch):
    plain, fast = _contexts(monkeypatch)
    try:
        saved = total = 0
        for n, seed in ((1000, 421), (2500, 5), (6000, 11)):
            D, z, a = gen_symarrow(n, seed=seed)
            r0, r1 = plain.symarrow(D, z, a), fast.symarrow(D, z, a)
            _same(r0, r1, ("symarrow", n))
            saved += int(r0.steps.sum()) - int(r1.steps.sum()); total += int(r0.steps.sum())
            D, u, rho = gen_symdpr1(n, seed=seed)
            _same(plain.symdpr1(D, u, rho), fast.symdpr1(D, u, rho), ("symdpr1", n))
            for tip in (1, 0):
                D, z = gen_halfarrow(n - 1, tip, seed=seed)
                _same(plain.halfarrow(D, z), fast.halfarrow(D, z), ("halfarrow", n, tip))
        assert saved > 0.5 * total           # more than half of the sweeps are gone
        rng = np.random.default_rng(7)
        for N in (60, 700, 3000):
            for name, D, z, a in arrow_families(rng, N):
                _same(plain.symarrow(D, z, a), fast.symarrow(D, z, a), ("symarrow", name, N))
                rho = abs(a) + 0.1
                _same(plain.symdpr1(D, z, rho), fast.symdpr1(D, z, rho), ("symdpr1", name, N))
        rng = np.random.default_rng(11)
        for nd in (40, 500, 2500):
            for name, D, z in half_families(rng, nd):
                for tip in (1, 0):
                    zz = z if tip else z[:nd]
                    _same(plain.halfarrow(D, zz), fast.halfarrow(D, zz), ("halfarrow", name, nd, tip))
    finally:
        plain.close(); fast.close()


@pytest.mark.timeout(600)
def test_accelerated_bisection_at_the_headline_size(monkeypatch):
    """n = 32768, every eigenvalue / Info field bitwise, eigenvector columns compared on the device"""
    import torch
    plain, fast = _contexts(monkeypatch)
    try:
        n = 32768
        D, z, a = gen_symarrow(n, seed=421)
        k0, k1 = 1, n
        outs = []
        for ctx in (plain, fast):
            U = torch.empty((8192, n), dtype=torch.float64, device="cuda")
            r = ctx.symarrow(D, z, a, vectors=False)
            part = ctx.resolve(12000, 12000 + 8191, vectors=False, U_dev=U.data_ptr(), ldu=n)
            torch.cuda.synchronize()
            outs.append((r, part, U))
        _same(outs[0][0], outs[1][0], "n=32768")
        _same(outs[0][1], outs[1][1], "n=32768 k-range")
        assert torch.equal(outs[0][2], outs[1][2])
        s0, s1 = int(outs[0][0].steps.sum()), int(outs[1][0].steps.sum())
        print(f"sweeps: plain {s0 / n:.1f} per eigenpair, accelerated {s1 / n:.1f}")
        assert s1 < 0.5 * s0
    finally:
        plain.close(); fast.close()


def _fuzz_problem(rng, N):
    """a random structured spectrum: uniform / clustered / graded pieces glued together, weights over many decades"""
    parts, left = [], N
    while left > 0:
        m = int(min(left, rng.integers(1, max(2, N // 2))))
        kind = rng.integers(0, 5)
        c = rng.normal() * 10.0 ** rng.integers(-3, 4)
        if kind == 0: p = c + rng.random(m)
        elif kind == 1: p = c + 10.0 ** rng.integers(-12, -2) * rng.random(m)            # a cluster
        elif kind == 2: p = c * 2.0 ** (-40 * rng.random(m))                             # graded towards 0
        elif kind == 3: p = c + np.arange(m) * 10.0 ** rng.integers(-9, 0)               # equispaced
        else: p = rng.standard_normal(m) * 10.0 ** rng.integers(-6, 7)
        parts.append(p); left -= m
    D = np.unique(np.concatenate(parts))[::-1].copy()
    D = D[np.isfinite(D)]
    keep = np.concatenate([[True], np.diff(D) < 0])
    D = D[keep]
    w = (rng.random(D.size) + 1e-3) * 10.0 ** (rng.integers(-12, 3) * rng.random(D.size))
    w *= np.where(rng.random(D.size) < 0.5, 1.0, -1.0)
    return D, w


@pytest.mark.timeout(900)
def test_accelerated_bisection_fuzz(monkeypatch):
    """random structured inputs, all three matrix types: every output bit equals plain bisection's"""
    plain, fast = _contexts(monkeypatch)
    try:
        rng = np.random.default_rng(20261020)
        done = 0
        for case in range(90):
            N = int(rng.choice([7, 33, 120, 500, 1100, 2300]))
            D, w = _fuzz_problem(rng, N)
            if D.size < 3:
                continue
            a = float(rng.normal() * 10.0 ** rng.integers(-3, 6))
            kind = case % 3
            if kind == 0:
                _same(plain.symarrow(D, w, a), fast.symarrow(D, w, a), ("fuzz symarrow", case, D.size))
            elif kind == 1:
                rho = abs(a) + 1e-3
                _same(plain.symdpr1(D, w, rho), fast.symdpr1(D, w, rho), ("fuzz symdpr1", case, D.size))
            else:
                Dh = np.unique(np.abs(D))[::-1].copy()
                Dh = Dh[Dh > 0]
                if Dh.size < 3:
                    continue
                tip = case % 2
                zh = np.resize(w, Dh.size + tip)
                _same(plain.halfarrow(Dh, zh), fast.halfarrow(Dh, zh), ("fuzz halfarrow", case, Dh.size, tip))
            done += 1
        assert done >= 80
    finally:
        plain.close(); fast.close()


@pytest.mark.timeout(900)
def test_accelerated_bisection_large_structured_families(monkeypatch):
    """the stress families at n = 12000 (several eigenpairs per slot, time slicing, Remedy-3 round through the ring):
    eigenvalues, Info, status and the eigenvectors (compared on the device) bit for bit"""
    import torch
    plain, fast = _contexts(monkeypatch)
    try:
        N = 12000
        U0 = torch.empty((N + 1, N + 1), dtype=torch.float64, device="cuda")
        U1 = torch.empty((N + 1, N + 1), dtype=torch.float64, device="cuda")
        V0 = torch.empty((N + 1, N + 1), dtype=torch.float64, device="cuda")
        V1 = torch.empty((N + 1, N + 1), dtype=torch.float64, device="cuda")

        def both(call, what, nrows, ncols, v=False):
            U0.zero_(); U1.zero_()
            if v: V0.zero_(); V1.zero_()
            r0 = call(plain, U0, V0); r1 = call(fast, U1, V1)
            torch.cuda.synchronize()
            _same(r0, r1, what)
            assert torch.equal(U0.view(torch.int64), U1.view(torch.int64)), what        # bit patterns (NaN columns of flagged eigenpairs included)
            if v: assert torch.equal(V0.view(torch.int64), V1.view(torch.int64)), what
            return int(r0.steps.sum()), int(r1.steps.sum())

        rng = np.random.default_rng(7)
        s0 = s1 = 0
        for name, D, z, a in arrow_families(rng, N):
            if np.any(np.diff(D) >= 0):
                D = np.unique(D)[::-1].copy(); z = z[:D.size]
            n = D.size + 1
            a0, a1 = both(lambda c, U, V: c.symarrow(D, z, a, vectors=False, U_dev=U.data_ptr(), ldu=n), ("symarrow", name), n, n)
            s0 += a0; s1 += a1
            rho = abs(a) + 0.1
            both(lambda c, U, V: c.symdpr1(D, z, rho, vectors=False, U_dev=U.data_ptr(), ldu=D.size), ("symdpr1", name), D.size, D.size)
        rng = np.random.default_rng(11)
        for name, D, z in half_families(rng, N - 1):
            for tip in (1, 0):
                zz = z if tip else z[:D.size]
                both(lambda c, U, V: c.halfarrow(D, zz, vectors=False, U_dev=U.data_ptr(), ldu=N + 1, V_dev=V.data_ptr(), ldv=N + 1),
                     ("halfarrow", name, tip), N + 1, N + 1, v=True)
        assert s1 < 0.6 * s0
    finally:
        plain.close(); fast.close()
