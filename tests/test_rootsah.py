"""rootsah (src/arrowhead7.jl:53-120): polynomial roots through the arrowhead companion
matrix with a double-double border.  CPU: the host-side companion construction + the oracle's restatement of the per-k
solver are pinned against the reference's known answers (test/runtests.jl:119-160: Wilkinson p18, Example 2).
GPU: the eigenvalue-only kernel (ah_symarrow_eigvals_dd) against the oracle and the same known answers."""
import numpy as np
import pytest

import arrowhead_b200 as ab
import oracle

EPS = float(np.finfo(np.float64).eps)
TAU = np.array([1e2, 1e2, 1e2, 1e2, 1e2])           # rootsah's default

P18 = [6402373705728000, -22376988058521600, 34012249593822720, -30321254007719424, 17950712280921504,
       -7551527592063024, 2353125040549984, -557921681547048, 102417740732658, -14710753408923, 1661573386473,
       -147560703732, 10246937272, -549789282, 22323822, -662796, 13566, -171, 1]              # runtests.jl:119-137
E5 = [-618970019642690000010608640, 4181389724724490601097907890741292883247104,
      -6277101735386680066937501969125693243111159424202737451008, 713623846352979940529142984724747568191373312,
      -20282409603651670423947251286016, 1]                                                    # runtests.jl:143-148
E5_ROOTS = [2.028240960365167e31, 1.759218623050247e13, 1.7592185858329531e13, 4.440892098500623e-16,
            2.2204460492503136e-16]                                                             # runtests.jl:156-160


def derivative_roots(coeffs):
    """D = roots(derivative(p)) to full double precision (the reference uses Polynomials.roots / rootsWDK)."""
    import mpmath as mp
    mp.mp.prec = 400
    n = len(coeffs) - 1
    d = [mp.mpf(int(c) if float(c).is_integer() else c) * k for k, c in enumerate(coeffs)][1:]
    r = mp.polyroots(d[::-1], maxsteps=500, extraprec=800)
    return np.array(sorted((float(mp.re(x)) for x in r), reverse=True))


def _oracle_roots(coeffs, D):
    Dm, zh, zl, ah, al = ab.companion_arrow(coeffs, D)
    z = zl + zh; zlo = (zh - z) + zl
    a = al + ah; alo = (ah - a) + al
    return oracle.get_lib().rootsah_eigvals(Dm, z, zlo, a, alo, TAU), (Dm, z, zlo, a, alo)


def test_double_double_helpers_match_the_oracle_c_code():
    from arrowhead_b200 import roots as R
    rng = np.random.default_rng(0)
    lib = oracle.get_lib()
    for op, fn in ((0, R.dd_add), (1, R.dd_sub), (2, R.dd_mul), (3, R.dd_div)):
        for _ in range(200):
            xh, yh = rng.standard_normal(2) * 10.0 ** rng.integers(-5, 6, 2)
            x = np.array([xh, xh * EPS * rng.random() * 0.4]); y = np.array([yh, yh * EPS * rng.random() * 0.4])
            out = np.zeros(2)
            lib.lib.ah_oracle_dd_op(op, x.ctypes.data_as(oracle.lib._dp), y.ctypes.data_as(oracle.lib._dp), out.ctypes.data_as(oracle.lib._dp))
            h, l = fn((np.float64(x[0]), np.float64(x[1])), (np.float64(y[0]), np.float64(y[1])))
            assert float(h) == out[0] and float(l) == out[1], (op, x, y)


def test_oracle_rootsah_wilkinson_p18():
    R, _ = _oracle_roots(P18, derivative_roots(P18))
    assert np.linalg.norm(R.lam - np.arange(18.0, 0.0, -1.0)) < 5 * EPS                        # runtests.jl:139


def test_oracle_rootsah_example2():
    R, _ = _oracle_roots([float(c) for c in E5], derivative_roots([float(c) for c in E5]))
    # runtests.jl:161 asks for norm(sort(R)-sort(Rtrue)) < 5 eps ABSOLUTE, i.e. bit-identical 1.76e13 roots; that needs
    # the reference's own barycentric points (rootsWDK(pd,r,1), BigFloat, out of scope).  With correctly rounded roots of
    # the derivative as D every root agrees with the reference's value to <= 2 ulp, the two tiny ones exactly.
    got, want = np.sort(R.lam), np.sort(E5_ROOTS)
    assert np.all(np.abs(got - want) <= 2 * EPS * np.abs(want))
    assert got[0] == want[0] and abs(got[1] - want[1]) <= 1e-30 and got[4] == want[4]


@pytest.mark.gpu
def test_gpu_rootsah_known_answers_and_oracle(gpu_ctx):
    roots, qout, res = ab.rootsah(P18, derivative_roots(P18), ctx=gpu_ctx, return_info=True)
    assert np.linalg.norm(roots - np.arange(18.0, 0.0, -1.0)) < 5 * EPS
    O, _ = _oracle_roots(P18, derivative_roots(P18))
    assert np.array_equal(res.sind, O.sind) and np.array_equal(qout, O.qout) and np.array_equal(res.status, O.status)
    assert np.all(np.abs(roots - O.lam) <= 10 * EPS * np.abs(O.lam))
    c5 = [float(c) for c in E5]
    roots5, q5 = ab.rootsah(c5, derivative_roots(c5), ctx=gpu_ctx)
    assert np.all(np.abs(np.sort(roots5) - np.sort(E5_ROOTS)) <= 2 * EPS * np.abs(np.sort(E5_ROOTS)))


@pytest.mark.gpu
def test_gpu_rootsah_degree_40_against_the_oracle(gpu_ctx):
    """a larger case in the spirit of runtests.jl:163-179 (that one uses the BigFloat rootsah, out of scope): the roots
    x_j = cos((2j-1)pi/(2n)) with barycentric points between them; coefficients built in exact rational arithmetic."""
    from fractions import Fraction
    n = 40
    xs = [Fraction(j, n + 1) * 2 - 1 for j in range(1, n + 1)]          # exactly representable-ish distinct roots in (-1,1)
    c = [Fraction(1)]
    for x in xs:                                                         # prod (t - x)
        c = [Fraction(0)] + c
        for i in range(len(c) - 1):
            c[i] -= x * c[i + 1]
    coeffs = [float(v) for v in c]
    Dpts = derivative_roots(coeffs)
    roots, qout, res = ab.rootsah(coeffs, Dpts, ctx=gpu_ctx, return_info=True)
    Dm, zh, zl, ah, al = ab.companion_arrow(coeffs, Dpts)
    z = zl + zh; zlo = (zh - z) + zl; a = al + ah; alo = (ah - a) + al
    O = oracle.get_lib().rootsah_eigvals(Dm, z, zlo, a, alo, TAU)
    assert np.array_equal(res.sind, O.sind) and np.array_equal(qout, O.qout)
    assert np.all(np.abs(roots - O.lam) <= (10 + 0.1 * O.knu) * EPS * np.maximum(np.abs(O.lam), 1e-300))
