"""The library's own FP64 GEMM (csrc/ah_gemm.cuh, DMMA.8x8x4) behind ah_dgemm_batched: the back-transform of tdc()
(arrowhead7.jl:32-33).  Compared with a plain float64 reference (numpy matmul on the host) at shapes that exercise both
tile sizes, ragged edges, odd leading dimensions and sub-blocks that start at odd element offsets (8-byte alignment only);
integer-valued matrices must come out exactly."""
import numpy as np
import pytest

import arrowhead_b200 as ab
from helpers import EPS

pytestmark = pytest.mark.gpu


def _run(ctx, A_list, B_list, pad=3, off=1):
    """C_b = A_b B_b for equally shaped problems stored as sub-blocks of larger column-major device matrices."""
    import torch
    nb = len(A_list)
    m, k = A_list[0].shape
    n = B_list[0].shape[1]
    lda, ldb, ldc = m + pad, k + pad + 1, m + pad + 2
    bufA = torch.zeros((nb, k + 1, lda), dtype=torch.float64, device="cuda")      # [b][col][row]: column-major blocks
    bufB = torch.zeros((nb, n + 1, ldb), dtype=torch.float64, device="cuda")
    bufC = torch.full((nb, n + 1, ldc), -7.0, dtype=torch.float64, device="cuda")
    for b in range(nb):
        bufA[b, :k, off:off + m] = torch.from_numpy(np.ascontiguousarray(A_list[b].T)).cuda()
        bufB[b, :n, off:off + k] = torch.from_numpy(np.ascontiguousarray(B_list[b].T)).cuda()
    ptr = lambda buf, b: buf[b].data_ptr() + 8 * off
    ctx.dgemm_batched(m, n, k, [ptr(bufA, b) for b in range(nb)], lda, [ptr(bufB, b) for b in range(nb)], ldb,
                      [ptr(bufC, b) for b in range(nb)], ldc)
    torch.cuda.synchronize()
    out = bufC.cpu().numpy()
    C = [out[b, :n, off:off + m].T.copy() for b in range(nb)]
    for b in range(nb):                                  # nothing outside the m x n block was written
        mask = np.ones_like(out[b], bool); mask[:n, off:off + m] = False
        assert np.all(out[b][mask] == -7.0)
    return C


@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (3, 5, 2, 7), (4, 64, 64, 64), (2, 65, 130, 67), (1, 128, 128, 16), (1, 129, 257, 33),
                                   (3, 127, 255, 127), (1, 300, 301, 299), (1, 1024, 2049, 1024), (40, 31, 63, 31), (1, 513, 96, 515)])
def test_dgemm_matches_float64_reference(gpu_ctx, shape):
    nb, m, n, k = shape
    rng = np.random.default_rng(m * 7 + n)
    A = [rng.standard_normal((m, k)) for _ in range(nb)]
    B = [rng.standard_normal((k, n)) for _ in range(nb)]
    C = _run(gpu_ctx, A, B)
    for b in range(nb):
        ref = A[b] @ B[b]
        bound = (k + 4) * EPS * (np.abs(A[b]) @ np.abs(B[b]))          # summation-order independent forward error bound
        assert np.all(np.abs(C[b] - ref) <= bound), (shape, np.max(np.abs(C[b] - ref) / bound))


def test_dgemm_is_exact_on_integers(gpu_ctx):
    rng = np.random.default_rng(3)
    A = [rng.integers(-8, 9, (97, 45)).astype(np.float64)]
    B = [rng.integers(-8, 9, (45, 201)).astype(np.float64)]
    C = _run(gpu_ctx, A, B, pad=0, off=0)
    assert np.array_equal(C[0], A[0] @ B[0])


def test_dgemm_orthogonal_back_transform(gpu_ctx):
    """the use in tdc: an orthogonal E1.vectors times the rows of the merged eigenvector matrix stays orthonormal"""
    rng = np.random.default_rng(5)
    k = 384
    Q1, _ = np.linalg.qr(rng.standard_normal((k, k)))
    W, _ = np.linalg.qr(rng.standard_normal((2 * k + 1, 2 * k + 1)))
    C = _run(gpu_ctx, [Q1], [W[:k, :]])[0]
    full = np.vstack([C, W[k:, :]])
    assert np.linalg.norm(full.T @ full - np.eye(2 * k + 1)) <= 50 * k * EPS


def test_async_utilities_defer_the_synchronisation(gpu_ctx):
    import torch
    n = 64
    Ud = ab.DeviceMatrix(gpu_ctx, n, n)
    gpu_ctx.set_async(True)
    try:
        for _ in range(50):                                              # host arrays may be reused/freed right after each call
            gpu_ctx.set_identity(Ud.ptr, n, n, Ud.ld)
            rows = np.random.default_rng(0).permutation(n); cols = np.arange(n)
            Od = ab.DeviceMatrix(gpu_ctx, n, n)
            gpu_ctx.gather_permuted(Od.ptr, Od.ld, Ud.ptr, Ud.ld, rows, cols)
            rows[:] = 0
            gpu_ctx.wait()
            assert np.array_equal(Od.to_host(), np.eye(n)[np.random.default_rng(0).permutation(n)])
            Od.free()
    finally:
        gpu_ctx.set_async(False)
        gpu_ctx.wait()
