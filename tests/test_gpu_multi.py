"""ah_multi: several GPUs (or several contexts on one GPU) driven through the C ABI from ONE host thread, and the NCCL
all-gather of U inside the library (VERDICT r1 "missing" item 2).

On a one-GPU box the sharding logic is exercised with two and three contexts on device 0: results must be bit-identical
to the single-context solve (the k-range never changes an eigenpair's arithmetic).  The NCCL part needs two devices
(NCCL refuses two ranks on one GPU) and is skipped otherwise; `gpurun --gpus 2` runs it."""
import numpy as np
import pytest

import arrowhead_b200 as ab
from helpers import gen_halfarrow, gen_symarrow, gen_symdpr1

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_k_partition_matches_the_python_sharding():
    lib = ab.load_library()
    import ctypes as C
    for nk, world in ((10, 3), (7, 8), (32768, 8), (1, 1), (16385, 4)):
        for r in range(world):
            k0, k1 = C.c_int64(), C.c_int64()
            lib.ah_k_partition(nk, world, r, C.byref(k0), C.byref(k1))
            assert (k0.value, k1.value) == ab.k_partition(nk, world, r)


@pytest.mark.parametrize("ndev", [1, 2, 3])
def test_multi_host_vectors_bitwise_equal_single_context(gpu_ctx, ndev):
    m = ab.MultiContext(devices=[0] * ndev)
    try:
        n = 1500
        D, z, a = gen_symarrow(n, seed=5)
        ref = gpu_ctx.symarrow(D, z, a)
        got = m.symarrow(D, z, a)
        for f in ("lam", "sind", "kb", "kz", "knu", "krho", "qout", "steps", "status"):
            assert np.array_equal(getattr(got, f), getattr(ref, f)), f
        assert np.array_equal(got.U, ref.U)
        assert len(got.stats["devices"]) == ndev and got.stats["elapsed_ms"] > 0
        assert sum(d["n_fast"] + d["n_full"] for d in got.stats["devices"]) == n
        D2, u2, rho = gen_symdpr1(900, seed=5)
        ref = gpu_ctx.symdpr1(D2, u2, rho); got = m.symdpr1(D2, u2, rho)
        assert np.array_equal(got.lam, ref.lam) and np.array_equal(got.U, ref.U) and np.array_equal(got.status, ref.status)
        D3, z3 = gen_halfarrow(700, 1, seed=5)
        ref = gpu_ctx.halfarrow(D3, z3); got = m.halfarrow(D3, z3)
        assert np.array_equal(got.lam, ref.lam) and np.array_equal(got.U, ref.U) and np.array_equal(got.V, ref.V)
        with pytest.raises(ab.ArrowheadError):                       # contract violations come back from every device
            m.symarrow(D[::-1].copy(), z, a)
    finally:
        m.close()


def test_multi_device_resident_full_buffers(gpu_ctx):
    """device-resident mode: every context writes its column block into its own full-size buffer at the block's offset"""
    import torch
    m = ab.MultiContext(devices=[0, 0])
    try:
        n = 1200
        D, z, a = gen_symarrow(n, seed=9)
        bufs = [torch.zeros((n, n), dtype=torch.float64, device="cuda:0") for _ in range(2)]      # row c = column c
        got = m.symarrow(D, z, a, vectors=False, U_full=[b.data_ptr() for b in bufs], ldu=n)
        torch.cuda.synchronize()
        ref = gpu_ctx.symarrow(D, z, a)
        assert np.array_equal(got.lam, ref.lam)
        for r in range(2):
            k0, k1 = m.partition(n, r)
            blk = bufs[r].cpu().numpy()
            assert np.array_equal(blk[k0 - 1:k1].T, ref.U[:, k0 - 1:k1])
            other = np.delete(blk, np.s_[k0 - 1:k1], axis=0)
            assert not other.any()                                   # nothing outside the own block was touched
    finally:
        m.close()


@pytest.mark.skipif(_ngpu() < 2, reason="NCCL all-gather needs two devices")
@pytest.mark.parametrize("n", [2048, 2051])                          # equal blocks (ncclAllGather) and ragged (ncclBroadcast per block)
def test_multi_allgather_columns_over_nccl(n):
    import torch
    ndev = min(_ngpu(), 8)
    m = ab.MultiContext(ndev=ndev)
    try:
        D, z, a = gen_symarrow(n, seed=13)
        bufs = [torch.zeros((n, n), dtype=torch.float64, device=f"cuda:{r}") for r in range(ndev)]
        got = m.symarrow(D, z, a, vectors=False, U_full=[b.data_ptr() for b in bufs], ldu=n)
        ms = m.allgather_columns([b.data_ptr() for b in bufs], n, n)
        assert ms > 0
        ref = ab.get_context(0).symarrow(D, z, a)
        assert np.array_equal(got.lam, ref.lam)
        for r in range(ndev):
            assert np.array_equal(bufs[r].cpu().numpy().T, ref.U), r
    finally:
        m.close()
