import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure, oracle/pyoracle.py)."""
    from oracle import pyoracle

    pyoracle.build()
    return pyoracle


# ---- "bit-exact" means bit patterns.  np.array_equal(a, b) is True for +0.0 vs -0.0, so an exact-zero output with the
# wrong sign passed every parity test until the end of round 2.  For float64 arrays the tests' array_equal
# therefore compares the int64 views (all NaNs count as one value: the kernels and the oracle fill failed instances with
# quiet NaNs whose payloads are not part of the spec).  NOC_TEST_VALUE_EQUALITY=1 restores numpy's semantics.
import numpy as _np

_value_array_equal = _np.array_equal


def _bits_array_equal(a, b, *args, **kw):
    if (isinstance(a, _np.ndarray) and isinstance(b, _np.ndarray) and a.dtype == _np.float64 and b.dtype == _np.float64
            and not args and not kw):
        if a.shape != b.shape:
            return False
        a, b = _np.ascontiguousarray(a), _np.ascontiguousarray(b)
        na, nb = _np.isnan(a), _np.isnan(b)
        if not _value_array_equal(na, nb):
            return False
        ia, ib = a.view(_np.int64), b.view(_np.int64)
        return bool(_np.all((ia == ib) | na))
    return _value_array_equal(a, b, *args, **kw)


@pytest.fixture(autouse=True, scope="session")
def _bitwise_array_equal():
    if os.environ.get("NOC_TEST_VALUE_EQUALITY"):
        yield
        return
    _np.array_equal = _bits_array_equal
    yield
    _np.array_equal = _value_array_equal
