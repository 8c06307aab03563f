"""Host-side logic of the package that needs no GPU: types, ordering, rotations, sharding arithmetic."""
import numpy as np
import pytest

import arrowhead_b200 as ab
import oracle


def test_types_match_reference_getindex():
    A = ab.SymArrow([1, 2, 3, 4], [1, 1, 1, 1], -3, 5)          # arrowhead1.jl docstring example
    M = A.dense()
    assert M.shape == (5, 5) and M[4, 4] == -3 and np.all(M[4, :4] == 1) and np.all(np.diag(M)[:4] == [1, 2, 3, 4])
    B = ab.SymArrow([1.0, 2.0, 3.0], [4.0, 5.0, 6.0], 7.0, 2)
    Md = B.dense()
    for i in range(4):
        for j in range(4):
            assert Md[i, j] == B[i, j]
    assert np.allclose(Md, Md.T)
    P = ab.SymDPR1([1, 2, 3, 4], [1, 1, 1, 1], 3)               # arrowhead1.jl docstring example
    assert np.array_equal(P.dense(), np.full((4, 4), 3.0) + np.diag([1, 2, 3, 4]))
    H = ab.HalfArrow([1, 2, 3, 4], [1, 1, 1, 1])
    assert H.shape == (4, 5)
    H2 = ab.HalfArrow([1, 2, 3, 4], [1, 1, 1, 1, 1])
    assert H2.shape == (5, 5) and H2.dense()[4, 4] == 1
    with pytest.raises(ValueError):
        ab.HalfArrow([1, 2], [1, 2, 3, 4])


def test_generators_shapes():
    rng = np.random.default_rng(0)
    A = ab.GenSymArrow(10, 10, rng)
    assert A.D.size == 9 and A.i == 10 and (A.D >= 0).all() and (A.D < 1).all()
    assert ab.GenSymDPR1(7, rng).u.size == 7
    assert ab.GenHalfArrow(5, 0, rng).shape == (4, 5) and ab.GenHalfArrow(5, 1, rng).shape == (5, 5)


def test_sortperm_matches_julia_semantics():
    x = np.array([1.0, 3.0, -0.0, 3.0, 0.0, 2.0])
    p = ab.jl_sortperm_desc(x)
    assert list(p) == [1, 3, 5, 0, 4, 2]                          # stable; 0.0 before -0.0
    assert np.array_equal(p, oracle.jl_sortperm_desc(x))


def test_givens_matches_oracle_and_zeroes_second_entry():
    rng = np.random.default_rng(2)
    for _ in range(50):
        f, g = rng.standard_normal(2)
        i1, i2, c, s, r = ab.givens(f, g, 3, 7)
        assert (i1, i2) == (3, 7)
        assert abs(c * f + s * g - r) < 1e-15 * max(1, abs(r)) and abs(-s * f + c * g) < 1e-15
        assert ab.givens(f, g, 3, 7) == oracle.givens(f, g, 3, 7)
    assert ab.givens(1.0, 0.0, 0, 1)[2:] == (1.0, 0.0, 1.0)
    assert ab.givens(0.0, 2.0, 0, 1)[2:] == (0.0, 1.0, 2.0)
    assert ab.givens(1.0, 2.0, 5, 2)[:2] == (2, 5)


@pytest.mark.parametrize("nk,world", [(32768, 8), (10, 4), (3, 8), (1, 1), (65536, 3)])
def test_k_partition_covers_everything_once(nk, world):
    seen = []
    sizes = []
    for r in range(world):
        k0, k1 = ab.k_partition(nk, world, r)
        sizes.append(max(0, k1 - k0 + 1))
        seen.extend(range(k0, k1 + 1))
    assert seen == list(range(1, nk + 1))
    assert max(sizes) - min(sizes) <= 1
