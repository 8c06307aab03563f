"""The oracle against the reference's own known answers (test/runtests.jl) and its property tests.

These pin the CPU restatement (oracle/) before it is trusted as the checker of the CUDA path."""
import numpy as np
import pytest

from helpers import EPS, gen_halfarrow, gen_symarrow, gen_symdpr1
from oracle import eigen_symarrow, eigen_symdpr1, svd_halfarrow, tdc


def rel_err(x, ref):
    return np.max(np.abs(x - ref) / np.abs(ref)) / EPS


@pytest.mark.parametrize("tag", ["symarrow_ex1", "symarrow_ex2", "symarrow_ex3"])
def test_symarrow_known_answers(golden, tag):
    c = golden[tag]
    E = eigen_symarrow(c["D"], c["z"], c["a"], c["i"])
    assert rel_err(E.values, c["values"]) < c["values_tol_eps"]          # runtests.jl:26,35,42
    if "v4" in c:
        assert rel_err(E.vectors[:, 3], c["v4"]) < c["v4_tol_eps"]       # runtests.jl:28


def test_symarrow_expected_branches(golden):
    """Branch counters listed in SURVEY.md Appendix B for the three paper examples."""
    E1 = eigen_symarrow(golden["symarrow_ex1"]["D"], golden["symarrow_ex1"]["z"], golden["symarrow_ex1"]["a"], 6)
    assert list(E1.info.sind) == [1, 1, 3, 3, 4, 5] and not E1.info.qout.any()
    assert E1.info.steps[0] > 100 and E1.info.steps[2] > 100          # Remedy 3 on k=1,3 (a second bisection)
    E2 = eigen_symarrow(golden["symarrow_ex2"]["D"], golden["symarrow_ex2"]["z"], golden["symarrow_ex2"]["a"], 5)
    assert list(E2.info.sind) == [1, 1, 2, 3, 4]
    E3 = eigen_symarrow(golden["symarrow_ex3"]["D"], golden["symarrow_ex3"]["z"], golden["symarrow_ex3"]["a"], 6)
    assert list(E3.info.sind) == [1, 2, 3, 4, 5, 5] and list(E3.info.qout) == [0, 1, 1, 1, 1, 1]
    assert abs(E3.info.kz[1] - 9999999999.666666) < 1e-5


@pytest.mark.parametrize("tag", ["symdpr1_ex1", "symdpr1_ex2"])
def test_symdpr1_known_answers(golden, tag):
    c = golden[tag]
    E = eigen_symdpr1(c["D"], c["u"], c["r"])
    assert np.max(np.abs(c["values"] - E.values) / np.abs(E.values)) / EPS < c["values_tol_eps"]   # runtests.jl:63,72
    if "v4" in c:
        assert rel_err(E.vectors[:, 3], c["v4"]) < c["v4_tol_eps"]                                  # runtests.jl:65


def test_symdpr1_gu_orthogonality(golden):
    c = golden["symdpr1_ex3_gu"]
    E = eigen_symdpr1(c["D"], c["u"], c["r"])
    U = E.vectors
    assert np.linalg.norm(U.T @ U - np.eye(4)) < c["orth_tol_eps"] * EPS                            # runtests.jl:78
    assert list(E.info.sind) == [1, 2, 2, 3] and list(E.info.qout) == [0, 1, 1, 1]


def test_tdc_wilkinson_w21(golden):
    c = golden["tdc_w21"]
    L, W = tdc(c["dv"], c["ev"])
    assert np.linalg.norm(np.sort(L) - np.sort(c["values"])) < c["norm_tol_eps"] * EPS              # runtests.jl:115
    # with the algebraically consistent rotation the vectors are eigenvectors too
    L2, W2 = tdc(c["dv"], c["ev"], rotation="transpose")
    T = np.diag(c["dv"]) + np.diag(c["ev"], 1) + np.diag(c["ev"], -1)
    assert np.linalg.norm(T @ W2 - W2 * L2) < 200 * EPS * np.linalg.norm(T)


def test_random_symarrow_residual(golden):
    """runtests.jl:15-19 on our own rand stream (Julia's is not reproducible here)."""
    rng = np.random.default_rng(421)
    for _ in range(5):
        D, z, a = rng.random(9), rng.random(9), rng.random()
        E = eigen_symarrow(D, z, a, 10)
        A = np.zeros((10, 10)); A[np.arange(9), np.arange(9)] = D; A[:9, 9] = z; A[9, :9] = z; A[9, 9] = a
        assert np.linalg.norm(A @ E.vectors - E.vectors * E.values) < golden["properties"]["symarrow_random_residual_tol_eps"] * EPS


def test_random_symdpr1_residual(golden):
    rng = np.random.default_rng(421)
    for _ in range(5):
        D, u, r = rng.random(10), rng.random(10), rng.random()
        E = eigen_symdpr1(D, u, r)
        A = np.diag(D) + r * np.outer(u, u)
        assert np.linalg.norm(A @ E.vectors - E.vectors * E.values) < golden["properties"]["symdpr1_random_residual_tol_eps"] * EPS


def test_halfarrow_graded_orthogonality(golden):
    """runtests.jl:83-87: 8x9 HalfArrow with entries exp(20(rand-0.5)); needs the double-double branch."""
    rng = np.random.default_rng(421)
    tol = golden["properties"]["halfarrow_orth_tol_eps"] * EPS
    used_dd = 0
    for _ in range(10):
        D = np.sort(np.exp(20 * (rng.random(8) - 0.5)))[::-1]
        z = np.exp(20 * (rng.random(8) - 0.5))
        S = svd_halfarrow(D, z)
        assert np.linalg.norm(S.U.T @ S.U - np.eye(8)) < tol
        assert np.linalg.norm(S.Vt.T.T @ S.Vt.T - np.eye(8)) < tol       # S.Vt*S.V
        A = np.zeros((8, 9)); A[np.arange(8), np.arange(8)] = D; A[:, 8] = z
        sv = np.linalg.svd(A, compute_uv=False)
        assert np.max(np.abs(S.S - sv)) < 20 * EPS * sv[0]      # LAPACK is only normwise accurate here
        used_dd += int(S.info.qout.any())
    assert used_dd > 0


@pytest.mark.parametrize("tip", [0, 1])
def test_halfarrow_matches_dense_svd(tip):
    D, z = gen_halfarrow(40, tip, seed=7)
    sgn = np.where(np.random.default_rng(3).random(40) < 0.5, -1.0, 1.0)
    Dsig = D * sgn
    rng = np.random.default_rng(11)
    p = rng.permutation(40)
    Din = Dsig[p]; zin = np.concatenate([z[:40][p], z[40:]])
    S = svd_halfarrow(Din, zin)
    A = np.zeros((40 + tip, 41)); A[np.arange(40), np.arange(40)] = Din; A[: zin.size, 40] = zin
    sv = np.linalg.svd(A, compute_uv=False)
    assert np.max(np.abs(S.S - sv[: S.S.size])) < 100 * EPS
    k = S.S.size
    R = A - S.U @ np.diag(S.S) @ S.Vt[:k, :]
    assert np.linalg.norm(R) < 200 * EPS


def test_oracle_medium_sizes_against_lapack():
    D, z, a = gen_symarrow(400, seed=5)
    E = eigen_symarrow(D, z, a, 400)
    A = np.zeros((400, 400)); A[np.arange(399), np.arange(399)] = D; A[:399, 399] = z; A[399, :399] = z; A[399, 399] = a
    w = np.linalg.eigvalsh(A)[::-1]
    assert np.max(np.abs(E.values - w)) < 100 * EPS
    assert np.linalg.norm(E.vectors.T @ E.vectors - np.eye(400)) < 400 * EPS
    assert np.linalg.norm(A @ E.vectors - E.vectors * E.values) < 400 * EPS * np.linalg.norm(A)
    assert (E.info.steps > 100).sum() > 0        # Remedy 3 fires on ~1 % of k for this distribution
    D, u, r = gen_symdpr1(300, seed=6)
    E = eigen_symdpr1(D, u, r)
    A = np.diag(D) + r * np.outer(u, u)
    assert np.max(np.abs(E.values - np.linalg.eigvalsh(A)[::-1])) < 100 * EPS * np.linalg.norm(A, 2)
    assert np.linalg.norm(A @ E.vectors - E.vectors * E.values) < 300 * EPS * np.linalg.norm(A)


def test_negative_rho_dpr1():
    D, u, r = gen_symdpr1(50, seed=9)
    E = eigen_symdpr1(D, u, -r)
    A = np.diag(D) - r * np.outer(u, u)
    assert np.max(np.abs(E.values - np.linalg.eigvalsh(A)[::-1])) < 50 * EPS
    assert np.linalg.norm(A @ E.vectors - E.vectors * E.values) < 100 * EPS


def test_double_double_ops_exact(oracle_lib):
    """Dekker operations of DoubleDouble.jl against exact rational arithmetic."""
    from fractions import Fraction
    rng = np.random.default_rng(1)
    for _ in range(200):
        x = (float(rng.standard_normal()), 0.0)
        y = (float(rng.standard_normal()) * 1e-3, 0.0)
        for op, f in (("add", lambda a, b: a + b), ("sub", lambda a, b: a - b), ("mul", lambda a, b: a * b),
                      ("div", lambda a, b: a / b)):
            hi, lo = oracle_lib.dd_op(op, x, y)
            exact = f(Fraction(x[0]), Fraction(y[0]))
            got = Fraction(hi) + Fraction(lo)
            assert abs(got - exact) <= abs(exact) * Fraction(2) ** -100


def test_deflation_paths():
    D = np.array([3.0, 2, 2, 1, 0.5, 0.5, 0.5]); z = np.array([1.0, 2, 3, 0, 1, 1, 1]); a = 0.3
    E = eigen_symarrow(D, z, a, 3, rotation="transpose")
    n = 8; t = 2
    idx = [j for j in range(n) if j != t]
    A = np.zeros((n, n)); A[idx, idx] = D; A[idx, t] = z; A[t, idx] = z; A[t, t] = a
    assert np.linalg.norm(A @ E.vectors - E.vectors * E.values) < 50 * EPS
    # the reference applies R instead of R' (arrowhead3.jl:635): eigenvalues agree, vectors stay orthonormal
    Er = eigen_symarrow(D, z, a, 3)
    assert np.allclose(Er.values, E.values, rtol=0, atol=0)
    assert np.linalg.norm(Er.vectors.T @ Er.vectors - np.eye(n)) < 50 * EPS
