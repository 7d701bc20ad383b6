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


@pytest.mark.gpu
def test_facade_host_layer_is_clean_under_asan_and_ubsan():
    """The C++ host layer and its test driver built with -fsanitize=address,undefined against the (uninstrumented)
    product library: every buffer the facade hands to the C-ABI is an exact-size std::vector, so an out-of-bounds host
    write by the library's copy-back, or UB in the header-only facade, is reported.  (compute-sanitizer is closed on the
    GPU pool; this is the host-side tool that is left.)"""
    _build()
    exe = EXE + "_san"
    src = os.path.join(ROOT, "tests", "cpp", "facade_test.cpp")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I",
           os.path.join(ROOT, "include"), src, "-o", exe, "-L", os.path.join(ROOT, "noc_b200"), "-lnoc_b200", "-L",
           os.path.join(ROOT, "oracle"), "-lnoc_oracle", "-Wl,-rpath," + os.path.join(ROOT, "noc_b200"),
           "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-Wl,-rpath,/usr/local/cuda/lib64", "-pthread"]
    subprocess.check_call(cmd)
    env = dict(os.environ, ASAN_OPTIONS="protect_shadow_gap=0:detect_leaks=0:abort_on_error=0",
               UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=env)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "FACADE OK" in out, out[-4000:]
    assert "ERROR: AddressSanitizer" not in out and "runtime error" not in out, out[-4000:]
