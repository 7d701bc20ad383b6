"""The C++ host layer (include/lqr_control/batched_lqr.hpp): compiles with plain g++ against the C-ABI, fails
loudly without a GPU (CPU test), and reproduces the oracle bit for bit on a GPU (gpu test)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "facade_test")


def _build():
    import __graft_entry__ as g
    from noc_b200 import capi
    from oracle import pyoracle
    if not os.path.exists(capi.LIB_PATH):
        g.build()
    pyoracle.build()
    src = os.path.join(ROOT, "tests", "cpp", "facade_test.cpp")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
           "-L", os.path.join(ROOT, "noc_b200"), "-lnoc_b200", "-L", os.path.join(ROOT, "oracle"), "-lnoc_oracle",
           "-Wl,-rpath," + os.path.join(ROOT, "noc_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
           "-Wl,-rpath,/usr/local/cuda/lib64", "-pthread"]
    subprocess.check_call(cmd)


def test_facade_compiles_and_fails_loudly_without_gpu():
    import torch
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "NO DEVICE" in r.stdout and "no CPU fallback" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_facade_matches_oracle_on_gpu():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "FACADE OK" in r.stdout, r.stdout + r.stderr


def test_jacobian_sparsity_predicates_cover_the_emitters():
    """quad_jac_nz_* / dual_jac_nz_* (the masks the fused sweeps prune with) == the entries the emitters write."""
    src = os.path.join(ROOT, "tests", "cpp", "sparsity_test.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "sparsity_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", src, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "SPARSITY OK" in r.stdout, r.stdout + r.stderr
