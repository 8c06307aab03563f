import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EPS = float(np.finfo(np.float64).eps)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _hex(v):
    return np.array([float.fromhex(x) for x in v])


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "runtests_known_answers.json"), encoding="utf-8") as f:
        raw = json.load(f)
    out = {}
    for tag, c in raw.items():
        d = dict(c)
        for key in ("D", "z", "u", "values", "v4", "dv", "ev"):
            if key in d:
                d[key] = _hex(d[key])
        for key in ("a", "r"):
            if key in d:
                d[key] = float.fromhex(d[key])
        out[tag] = d
    return out


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    return oracle.get_lib()


@pytest.fixture(scope="session")
def oracle_hi():
    """same algorithm, long-double accumulators (measures the reference's own summation noise)"""
    import oracle
    return oracle.get_lib_hi()


@pytest.fixture(scope="session")
def gpu_ctx():
    import arrowhead_b200 as ab
    ab.build_library()
    return ab.get_context(0)
