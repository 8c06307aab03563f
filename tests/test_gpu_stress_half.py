"""HalfArrow families through the C ABI against the oracle (both shapes): graded poles / border, clustered poles,
a tiny tip entry.  Where the reference's own scheme is noise (oracle NaN, K_rho >= 1e3 in a Remedy) the harness
says so explicitly (tests/helpers.py::compare_range); everywhere else the plain bars apply."""
import numpy as np
import pytest

import oracle
from helpers import EPS, compare_range, shift_margin_halfarrow

pytestmark = pytest.mark.gpu


def families(rng, nd):
    yield "uniform", np.sort(rng.random(nd) + 0.01)[::-1].copy(), rng.random(nd + 1) - 0.5
    yield "graded D 1e-8..1", np.sort(10.0 ** (-8 * rng.random(nd)))[::-1].copy(), rng.random(nd + 1) - 0.5
    yield "graded z", np.sort(rng.random(nd) + 0.01)[::-1].copy(), 10.0 ** (-8 * rng.random(nd + 1))
    yield "clustered D", np.sort(1.0 + 1e-9 * np.arange(nd) * (1 + rng.random(nd)))[::-1].copy(), rng.random(nd + 1)
    yield "both graded", np.sort(np.exp(20 * (rng.random(nd) - 0.5)))[::-1].copy(), np.exp(20 * (rng.random(nd + 1) - 0.5))
    yield "tiny tip", np.sort(rng.random(nd) + 0.01)[::-1].copy(), np.concatenate([rng.random(nd), [1e-12]])


@pytest.mark.parametrize("nd", [40, 500])
def test_halfarrow_families_match_the_oracle(gpu_ctx, oracle_lib, oracle_hi, nd):
    rng = np.random.default_rng(11)
    for name, D, z in families(rng, nd):
        for tip in (1, 0):
            zz = z if tip else z[:nd]
            tau = oracle.default_tau(nd)
            G = gpu_ctx.halfarrow(D, zz)
            out = compare_range(G, oracle_lib.halfarrow(D, zz, tau), oracle_hi.halfarrow(D, zz, tau), margin_fn=lambda k: shift_margin_halfarrow(D, zz, k),
                                vec_tol=1000.0, max_excluded=max(2, G.lam.size // 20), poles=D)
            # backward error of the singular values against LAPACK (absolute accuracy eps*||A||)
            A = np.zeros((max(nd, zz.size), nd + 1)); A[np.arange(nd), np.arange(nd)] = D; A[:zz.size, nd] = zz
            sv = np.linalg.svd(A, compute_uv=False)[:G.lam.size]
            good = ~np.isnan(G.lam)
            assert np.max(np.abs(G.lam[good] - sv[good])) <= 100 * EPS * sv[0], (name, tip, out)
