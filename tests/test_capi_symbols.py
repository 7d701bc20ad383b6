"""CPU-only: the product library loads and exports every symbol include/lqr_control/noc_b200.h declares.
No compute call is made (there is no GPU here and no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lqr_control", "noc_b200.h")


@pytest.fixture(scope="module")
def lib():
    from noc_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return capi.load_library()


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(noc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 14
    for nme in names:
        assert hasattr(lib, nme), f"{nme} declared in the header but not exported"


def test_python_export_list_matches_header():
    from noc_b200 import capi
    assert sorted(capi.EXPORTS) == _declared()


def test_problem_struct_layout(lib):
    from noc_b200 import capi
    p = capi.problem(capi.MODEL_QUAD12, 100)
    assert (p.n, p.m, p.N, p.model, p.layout) == (12, 4, 100, 0, 0)
    assert (p.max_iter, p.n_alpha) == (50, 11)
    assert p.dt == 0.02 and p.tol == 1e-8 and p.armijo == 1e-4 and p.reg0 == 0.0
    p = capi.problem(capi.MODEL_DUAL24, 7)
    assert (p.n, p.m, p.N) == (24, 8, 7)
    assert p.u_min == -float("inf") and p.u_max == float("inf")        # no input box by default
    assert p.reg_schedule == capi.REG_X10 and p.reserved_ == 0
    assert ctypes.sizeof(capi.Problem) == 8 * 4 + 6 * 8 + 2 * 4


def test_version_and_no_gpu_fails_loudly(lib):
    from noc_b200 import capi
    assert b"sm_100a" in lib.noc_version()
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(capi.NocError):
            capi.Solver(0)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under noc_b200/ (python or CUDA sources) may import, include, link
    or dlopen anything under oracle/, and the product library must not depend on libnoc_oracle."""
    import re
    import subprocess
    pkg = os.path.join(ROOT, "noc_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")) and f != "Makefile":
                continue
            text = open(os.path.join(dirpath, f)).read()
            code = re.sub(r"//[^\n]*|/\*.*?\*/", "", text, flags=re.S)       # comments cite the oracle's functions; code may not
            if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or re.search(r'#include\s+"[^"]*oracle', code) \
                    or re.search(r"noc_oracle_\w+\s*\(", code) or "lnoc_oracle" in code or "libnoc_oracle" in code:
                bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    from noc_b200 import capi
    if os.path.exists(capi.LIB_PATH):
        deps = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
        assert "oracle" not in deps


def test_binding_checks_dtype_and_size_before_the_pointer_crosses():
    """the C-ABI takes raw pointers, so the ctypes mirror must refuse a float32 / int64 / too-small array
    instead of letting the library write fp64 / int32 data of its own size behind it."""
    import numpy as np
    from noc_b200 import capi
    assert capi._f(None, 10, "x") is None
    ok = np.zeros(12)
    assert capi._f(ok, 12, "x") == ok.ctypes.data
    with pytest.raises(TypeError):
        capi._f(np.zeros(12, dtype=np.float32), 12, "K0")
    with pytest.raises(TypeError):
        capi._i(np.zeros(4, dtype=np.int64), 4, "status")
    with pytest.raises(ValueError):
        capi._f(np.zeros(11), 12, "cost")
    with pytest.raises(ValueError):
        capi._f(np.zeros((4, 6))[:, ::2], 12, "x0")            # not contiguous
    import torch
    t = torch.zeros(12, dtype=torch.float64)
    assert capi._f(t, 12, "x") == t.data_ptr()
    with pytest.raises(TypeError):
        capi._f(torch.zeros(12, dtype=torch.float32), 12, "x")
    with pytest.raises(ValueError):
        capi._i(torch.zeros(3, dtype=torch.int32), 4, "iters")


def test_no_link_time_nccl_dependency():
    """The multi-GPU plane loads NCCL with dlopen at run time (inside a PyTorch process it shares torch's copy); the
    product library itself must stay loadable on a box without libnccl."""
    import subprocess
    from noc_b200 import capi
    deps = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "nccl" not in deps.lower(), deps


def test_time_parallel_chunk_plan_invariants(lib):
    """Host-side logic of NOC_FLAG_TIME_PARALLEL (no GPU needed): for every horizon the plan uses 1..8 chunks, every
    chunk has at least one step, the chunks tile the horizon, short horizons are not cut at all, and the last chunk —
    which only runs plain steps while the others also fold their interval elements — is the longest."""
    from noc_b200 import capi
    prev_chunks = 1
    for N in list(range(1, 130)) + [200, 333, 517, 1000, 4096]:
        chunks, ln = capi.plan_time_parallel(N)
        assert 1 <= chunks <= 8 and ln >= 1
        last = N - (chunks - 1) * ln
        assert last >= 1
        if chunks == 1:
            assert ln == N
        else:
            assert last >= ln                       # the cheap chunk is the long one
        if N <= 16:
            assert chunks == 1                      # a scan cannot beat sixteen plain steps
    assert capi.plan_time_parallel(100)[0] > 1 and capi.plan_time_parallel(200)[0] > 1
    with pytest.raises(capi.NocError):
        capi.plan_time_parallel(0)
