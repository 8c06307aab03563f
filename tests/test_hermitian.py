"""Hermitian wrappers (hermitian.jl:107-139): complex / quaternion border,
phase vector c = z./abs.(z), real solve, `Diagonal(c)*U`.  The reference has no test for them, so the CPU
test pins the restatement against numpy's dense Hermitian eigensolver and the GPU test compares with it."""
import numpy as np
import pytest

import oracle

EPS = float(np.finfo(np.float64).eps)


def _gen(n, ncomp, seed):
    rng = np.random.default_rng(seed)
    return rng.random(n - 1) - 0.3, rng.random((n - 1, ncomp)) - 0.5, float(rng.random())


@pytest.mark.parametrize("i", [1, 4, 12])
def test_oracle_hermarrow_complex_matches_dense_eigh(i):
    n = 12
    D, zc, a = _gen(n, 2, 5)
    E = oracle.eigen_hermarrow(D, zc, a, i)
    idx = [r for r in range(n) if r != i - 1]
    A = np.zeros((n, n), complex)
    A[idx, idx] = D
    A[idx, i - 1] = zc[:, 0] + 1j * zc[:, 1]
    A[i - 1, idx] = zc[:, 0] - 1j * zc[:, 1]
    A[i - 1, i - 1] = a
    w = np.linalg.eigvalsh(A)[::-1]
    assert np.max(np.abs(E.values - w)) <= 50 * EPS * np.abs(w).max()
    U = E.vectors
    assert np.linalg.norm(A @ U - U * E.values) <= 100 * EPS * np.linalg.norm(A)
    assert np.linalg.norm(U.conj().T @ U - np.eye(n)) <= 100 * EPS


def test_oracle_hermdpr1_complex_matches_dense_eigh():
    n = 9
    rng = np.random.default_rng(6)
    D, uc, r = rng.random(n), rng.random((n, 2)) - 0.5, -0.7
    E = oracle.eigen_hermdpr1(D, uc, r)
    u = uc[:, 0] + 1j * uc[:, 1]
    A = np.diag(D).astype(complex) + r * np.outer(u, u.conj())
    assert np.max(np.abs(E.values - np.linalg.eigvalsh(A)[::-1])) <= 50 * EPS
    assert np.linalg.norm(A @ E.vectors - E.vectors * E.values) <= 100 * EPS * np.linalg.norm(A)


def test_oracle_quaternion_phase_is_unit_and_real_border_is_modulus():
    D, zc, a = _gen(7, 4, 7)
    c, zr = oracle.drivers._phase_split(zc)
    assert np.allclose(np.sqrt((c ** 2).sum(axis=1)), 1.0, rtol=0, atol=4 * EPS)
    assert np.allclose(zr, np.sqrt((zc ** 2).sum(axis=1)), rtol=4 * EPS, atol=0)
    E = oracle.eigen_hermarrow(D, zc, a, 3)
    assert E.vectors.shape == (7, 7, 4)
    assert np.allclose(np.sqrt((E.vectors ** 2).sum(axis=2)), np.abs(oracle.eigen_symarrow(D, zr, a, 3).vectors), rtol=0, atol=8 * EPS)


@pytest.mark.gpu
@pytest.mark.parametrize("ncomp", [2, 4])
def test_gpu_hermarrow_matches_oracle(gpu_ctx, ncomp):
    import arrowhead_b200 as ab
    n, i = 300, 17
    D, zc, a = _gen(n, ncomp, 11)
    E, info = ab.eigen(ab.HermArrow(D, zc if ncomp == 4 else zc[:, 0] + 1j * zc[:, 1], a, i), ctx=gpu_ctx)
    O = oracle.eigen_hermarrow(D, zc, a, i)
    assert np.array_equal(info.sind, O.info.sind)
    assert np.all(np.abs(E.values - O.values) <= (10 + 0.1 * O.info.knu) * EPS * np.abs(O.values))
    G, R = np.asarray(E.vectors), np.asarray(O.vectors)
    assert G.shape == R.shape
    assert np.max(np.abs(G - R)) <= 1e4 * EPS                      # entries of unit vectors; Knu-limited like the real case
    if ncomp == 2:
        A = ab.HermArrow(D, zc[:, 0] + 1j * zc[:, 1], a, i).dense()
        assert np.linalg.norm(A @ G - G * E.values) <= n * EPS * np.linalg.norm(A)
        assert np.linalg.norm(G.conj().T @ G - np.eye(n)) <= n * EPS


@pytest.mark.gpu
def test_gpu_hermdpr1_matches_oracle_and_keeps_vectors_on_device(gpu_ctx):
    import arrowhead_b200 as ab
    n = 257
    rng = np.random.default_rng(13)
    D, uc, r = rng.random(n), rng.random((n, 2)) - 0.5, 0.9
    A = ab.HermDPR1(D, uc[:, 0] + 1j * uc[:, 1], r)
    E, info = ab.eigen(A, ctx=gpu_ctx)
    O = oracle.eigen_hermdpr1(D, uc, r)
    assert np.all(np.abs(E.values - O.values) <= (10 + 0.1 * O.info.knu) * EPS * np.abs(O.values))
    M = A.dense()
    assert np.linalg.norm(M @ E.vectors - E.vectors * E.values) <= n * EPS * np.linalg.norm(M)
    Ed, _ = ab.eigen(A, ctx=gpu_ctx, vectors="device")
    assert Ed.vectors.nrows == 2 * n and Ed.vectors.ncols == n      # interleaved ComplexF64 columns, resident
    Ed.vectors.free()
