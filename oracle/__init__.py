"""CPU oracle for the Arrowhead.jl all-eigenpairs drivers -- TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``arrowhead.jl_b200/``) imports this.  Allowed users: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

``arrowhead_oracle.c`` restates the per-eigenpair solvers (arrowhead3.jl:419-550, arrowhead4.jl:289-422,
arrowhead6.jl:344-508 and the inv/bisect kernels under them); ``drivers.py`` restates the host-side
drivers (sorting, deflation, back-permutation: arrowhead3.jl:567-657, arrowhead4.jl:438-551,
arrowhead6.jl:526-661, arrowhead7.jl:10-36).  The reference is Julia and cannot run in this image;
parity is pinned by the reference's own known-answer vectors (tests/golden/, tests/test_oracle_golden.py).
"""
from .lib import OracleLib, get_lib, get_lib_hi, build  # noqa: F401
from .drivers import (  # noqa: F401
    eigen_symarrow,
    eigen_symdpr1,
    eigen_hermarrow,
    eigen_hermdpr1,
    svd_halfarrow,
    tdc,
    default_tau,
    jl_sortperm_desc,
    givens,
)
