"""svd(::HalfArrow) with deflation: zeros in z and exact duplicates of |D| (arrowhead6.jl:569-632).

The reference's own branches cannot run as written, so there is no reference output to match; the
bar is the mathematics -- A = U*S*V', orthonormal factors, singular values equal to LAPACK's to 10 eps*||A|| -- for the
oracle driver (CPU) and for the product driver (GPU), and the two against each other."""
import numpy as np
import pytest

import oracle
from helpers import EPS, halfarrow_deflation_cases, halfarrow_dense

CASES = halfarrow_deflation_cases()


def _check_svd(name, D, z, U, S, Vt):
    A = halfarrow_dense(D, z)
    n0 = z.size
    assert U.shape == (n0, n0) and S.shape == (n0,) and Vt.shape == (n0, D.size + 1), name
    assert np.all(np.diff(S) <= 0) and np.all(S >= 0), name
    scale = max(1.0, np.linalg.norm(A))
    assert np.linalg.norm(A - U @ np.diag(S) @ Vt) <= 50 * n0 * EPS * scale, (name, np.linalg.norm(A - U @ np.diag(S) @ Vt))
    assert np.linalg.norm(U.T @ U - np.eye(n0)) <= 50 * n0 * EPS, name
    assert np.linalg.norm(Vt @ Vt.T - np.eye(n0)) <= 50 * n0 * EPS, name
    sv = np.linalg.svd(A, compute_uv=False)
    assert np.max(np.abs(S - sv[:n0])) <= 20 * EPS * scale, name


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_driver_deflation_is_a_valid_svd(case):
    name, D, z = case
    R = oracle.svd_halfarrow(D, z)
    _check_svd(name, D, z, R.U, R.S, R.Vt)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_product_driver_deflation_matches_the_oracle_driver(gpu_ctx, case):
    import arrowhead_b200 as ab
    name, D, z = case
    S, info = ab.svd(ab.HalfArrow(D, z), ctx=gpu_ctx)
    _check_svd(name, D, z, S.U, S.S, S.Vt)
    R = oracle.svd_halfarrow(D, z)
    assert np.max(np.abs(S.S - R.S)) <= 10 * EPS * max(1.0, R.S.max()), name
    assert np.array_equal(info.sind, R.info.sind) and np.array_equal(info.qout, R.info.qout), name
    # vectors: compare where the singular value is simple (duplicates span a plane: any orthonormal basis is valid)
    simple = np.ones(S.S.size, bool)
    simple[1:] &= np.abs(np.diff(R.S)) > 1e-8
    simple[:-1] &= np.abs(np.diff(R.S)) > 1e-8
    assert np.max(np.abs(S.U[:, simple] - R.U[:, simple]), initial=0.0) <= 1000 * EPS, name
    assert np.max(np.abs(S.Vt[simple] - R.Vt[simple]), initial=0.0) <= 1000 * EPS, name
    Sd, _ = ab.svd(ab.HalfArrow(D, z), ctx=gpu_ctx, vectors="device")
    assert np.array_equal(Sd.S, S.S)


@pytest.mark.gpu
def test_zero_tip_or_zero_pole_is_refused(gpu_ctx):
    import arrowhead_b200 as ab
    with pytest.raises(ValueError):
        ab.svd(ab.HalfArrow([2.0, 1.0], [1.0, 1.0, 0.0]), ctx=gpu_ctx)
    with pytest.raises(ValueError):
        ab.svd(ab.HalfArrow([2.0, 0.0], [1.0, 1.0]), ctx=gpu_ctx)
