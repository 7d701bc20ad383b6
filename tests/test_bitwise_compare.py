"""The parity tests' `np.array_equal` compares bit patterns for float64 arrays (tests/conftest.py)."""
import os

import numpy as np
import pytest


@pytest.mark.skipif(bool(os.environ.get("NOC_TEST_VALUE_EQUALITY")), reason="value equality requested")
def test_array_equal_sees_the_sign_of_zero():
    assert not np.array_equal(np.array([0.0, 1.0]), np.array([-0.0, 1.0]))
    assert np.array_equal(np.array([-0.0, 1.0]), np.array([-0.0, 1.0]))
    assert np.array_equal(np.array([np.nan, 2.0]), np.array([np.nan, 2.0]))        # NaN fills of failed instances
    assert not np.array_equal(np.array([np.nan, 2.0]), np.array([1.0, 2.0]))
    assert not np.array_equal(np.zeros(3), np.zeros(4))
    assert np.array_equal(np.array([1, 2], dtype=np.int32), np.array([1, 2], dtype=np.int32))
