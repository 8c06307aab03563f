"""Structured matrix families through the C ABI against the oracle: clustered poles, wide dynamic range in z,
huge arrow top, one-signed spectra, geometric poles.  These exercise Remedy 3 (single and repeated re-shift),
the double-double hand-over and the interlacing safeguard."""
import numpy as np
import pytest

import arrowhead_b200 as ab
import oracle
from helpers import EPS, compare_range, shift_margin_symarrow, shift_margin_symdpr1

pytestmark = pytest.mark.gpu


def families(rng, N):
    yield "clustered pairs 1e-10", np.sort(np.repeat(rng.random(N // 2), 2) + np.tile([0, 1e-10], N // 2))[::-1].copy(), rng.random(N), rng.random()
    yield "wide z range", np.sort(rng.random(N) * 2 - 1)[::-1].copy(), 10.0 ** (-8 * rng.random(N)), rng.random()
    yield "huge a", np.sort(rng.random(N))[::-1].copy(), rng.random(N), 1e12
    yield "all negative", np.sort(-rng.random(N) - 0.1)[::-1].copy(), rng.random(N) - 0.5, -3.0
    yield "tiny z 1e-100", np.sort(rng.random(N))[::-1].copy(), 1e-100 * (0.5 + rng.random(N)), rng.random()
    yield "mixed tiny z", np.sort(rng.random(N))[::-1].copy(), np.where(rng.random(N) < 0.3, 1e-30, 1.0) * (0.5 + rng.random(N)), rng.random()
    yield "geometric D", np.sort(2.0 ** (-np.arange(N) * 60.0 / N))[::-1].copy(), rng.random(N), 0.0
    yield "sym about 0", np.sort(np.concatenate([np.linspace(0.1, 1, N // 2), -np.linspace(0.1, 1, N - N // 2) * (1 + 1e-3)]))[::-1].copy(), rng.random(N), 0.0


@pytest.mark.parametrize("N", [60, 700])
def test_structured_families_match_the_oracle(gpu_ctx, oracle_lib, oracle_hi, N):
    rng = np.random.default_rng(7)
    seen = 0
    for name, D, z, a in families(rng, N):
        tau = oracle.default_tau(N)
        G = gpu_ctx.symarrow(D, z, a)
        compare_range(G, oracle_lib.symarrow(D, z, a, tau), oracle_hi.symarrow(D, z, a, tau),
                      margin_fn=lambda k: shift_margin_symarrow(D, z, a, k), max_excluded=max(2, N // 20), poles=D)
        flagged = (G.status & ab.ST_INTERLACE) != 0                                       # never silent: in bounds, or flagged
        assert np.all((G.lam[1:] <= D) | flagged[1:]) and np.all((G.lam[:-1] >= D) | flagged[:-1]), name
        assert flagged.sum() <= 1, name
        rho = abs(a) + 0.1
        G = gpu_ctx.symdpr1(D, z, rho)
        compare_range(G, oracle_lib.symdpr1(D, z, rho, tau), oracle_hi.symdpr1(D, z, rho, tau),
                      margin_fn=lambda k: shift_margin_symdpr1(D, z, rho, k), max_excluded=max(2, N // 20))
        flagged = (G.status & ab.ST_INTERLACE) != 0
        assert np.all((G.lam >= D) | flagged) and np.all((G.lam[1:] <= D[:-1]) | flagged[1:]), name
        seen += 1
    assert seen == 8


def test_interlacing_safeguard_catches_the_wrong_end_of_the_spectrum(gpu_ctx):
    """Geometric poles down to 2^-60, a = 0, k = N+1: the first-stage K_nu is ~1e19, nu is rounding noise, and the
    reference's Remedy 3 can re-shift past the eigenvalue and return the LARGEST eigenvalue for the smallest one
    (seen with this very input before the safeguard).  Now: correct value, or AH_ST_INTERLACE -- never silent."""
    import mpmath as mp
    N = 60
    rng = np.random.default_rng(7)
    for name, D, z, a in families(rng, N):
        if name == "geometric D":
            break
    G = gpu_ctx.symarrow(D, z, a)
    A = np.zeros((N + 1, N + 1)); A[np.arange(N), np.arange(N)] = D; A[:N, N] = z; A[N, :N] = z; A[N, N] = a
    mp.mp.prec = 200
    lam_min = float(min(mp.eigsy(mp.matrix(A.tolist()), eigvals_only=True)))
    assert (G.status[-1] & ab.ST_INTERLACE) or abs(G.lam[-1] - lam_min) <= 100 * EPS * abs(lam_min)
    assert G.lam[-1] < 0 or (G.status[-1] & ab.ST_INTERLACE)
