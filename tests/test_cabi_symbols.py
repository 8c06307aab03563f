"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the header declares,
and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    import arrowhead_b200 as ab
    ab.build_library()
    return ab.load_library()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "arrowhead_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ah_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    import arrowhead_b200 as ab
    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/arrowhead_b200.h but not exported"
    assert sorted(ab.EXPORTS) == names
    assert lib.ah_version() == 100


def test_no_cpu_fallback(lib):
    import torch
    import arrowhead_b200 as ab
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    with pytest.raises(ab.ArrowheadError) as ei:
        ab.Context(0)
    assert ei.value.code == -4                      # AH_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.ah_last_error(None)
    with pytest.raises(ab.ArrowheadError):
        ab.eigen(ab.SymArrow([2.0, 1.0], [1.0, 1.0], 0.5, 3))


def test_null_context_is_rejected(lib):
    D = np.array([2.0, 1.0]); z = np.array([1.0, 1.0])
    from arrowhead_b200._lib import AhResult, _dp
    r = AhResult()
    rc = lib.ah_symarrow_eigen(None, 2, D.ctypes.data_as(_dp), z.ctypes.data_as(_dp), 0.5, None, 1, 3, C.byref(r))
    assert rc == -1
    assert lib.ah_set_mode(None, 0) == -1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "arrowhead.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "oracle" not in src.lower() or f == "__init__.py" and "oracle" not in src, \
                    f"{f} mentions the oracle: the product path must not depend on it"
