"""The branch-free reciprocal used for the L D L^T pivots (riccati12.cuh: rcp_rn_normal) must equal the IEEE
correctly rounded 1/x bit for bit on its whole contract range (1e-280, 1e280) — the CPU oracle divides."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_rcp_bitexact_against_ieee_division():
    from noc_b200 import capi
    s = capi.Solver(0)
    rng = np.random.default_rng(3)
    n = 1 << 22
    mant = rng.uniform(1.0, 2.0, n)
    expo = rng.integers(-900, 900, n)
    x = np.ldexp(mant, expo)
    # hard mantissas: all ones, one, one + ulp, near sqrt(2), powers of two, typical pivots
    special = np.array([1.0, np.nextafter(1.0, 2.0), np.nextafter(2.0, 1.0), np.sqrt(2.0), 1.5, 3.0, 0.1, 0.3,
                        1e-279, 9.9e279, 2.0 ** -500, 2.0 ** 500, 7.0, 1.0 / 3.0])
    x[:special.size] = special
    x[special.size:2 * special.size] = np.nextafter(special, 0.0)
    x[1000:2000] = np.nextafter(2.0, 1.0) * np.ldexp(1.0, rng.integers(-900, 900, 1000))
    x[2000:200000] = rng.uniform(1e-3, 1e4, 198000)       # where real pivots live
    fast, ieee = s.selftest_rcp(x)
    assert np.array_equal(ieee, 1.0 / x)                   # device IEEE division == host division
    bad = np.flatnonzero(fast != ieee)
    assert bad.size == 0, (bad[:10], x[bad[:10]], fast[bad[:10]], ieee[bad[:10]])
    s.close()
