"""Randomised SLQ parity sweep (tools/fuzz_slq.py) as a test: random models, horizons, batches, input boxes, regularisation
schedules, indefinite R, aggressive starts, kernel families and schedules — bit-exact against the oracle in every case."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [3, 4])
def test_slq_fuzz_bitexact(seed):
    from tools import fuzz_slq
    out = fuzz_slq.run(cases=120, seed=seed, verbose=False)
    assert out["mismatches"] == 0
    cov = out["covered"]
    assert cov["backtracked"] > 0 and cov["boxed"] > 0 and cov["adaptive"] > 0 and len(cov["status"]) >= 3
