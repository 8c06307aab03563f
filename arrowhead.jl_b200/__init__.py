"""arrowhead_b200 -- B200-native all-eigenpairs drivers with the API of ivanslapnicar/Arrowhead.jl.

    from arrowhead_b200 import SymArrow, eigen
    E, info = eigen(SymArrow(D, z, a, i))            # E.values, E.vectors, info.sind/kb/kz/knu/krho/qout

The package directory is `arrowhead.jl_b200/` (not an importable name); `arrowhead_b200.py` at the repo
root aliases it.  Compute runs only in libarrowhead_b200.so (hand-written sm_100a CUDA, C ABI in
include/arrowhead_b200.h); there is no CPU fallback.
"""
from .types import (SVD, Eigen, GenHalfArrow, GenHermArrow, GenHermDPR1, GenSymArrow, GenSymDPR1, HalfArrow,  # noqa: F401
                    HermArrow, HermDPR1, Info, SymArrow,
                    SymDPR1)
from ._lib import (LIB_PATH, ArrowheadError, Context, MultiContext, get_context, load_library, EXPORTS,  # noqa: F401
                   ST_NU_INF, ST_NEEDS_BIGFLOAT, ST_POLE_RETURN, ST_REF_THROWS, ST_ITER_LIMIT, ST_INTERLACE)
from .drivers import (DeviceMatrix, eigen, eigen_hermarrow, eigen_hermdpr1, eigen_symarrow, eigen_symdpr1, givens, jl_sortperm_desc,  # noqa: F401
                      svd, svd_halfarrow, tdc, tdc_device, tdc_recursive)
from .roots import companion_arrow, rootsah  # noqa: F401
from .sharding import k_partition, solve_sharded  # noqa: F401
from .build import build_library  # noqa: F401

__version__ = "0.1.0"
