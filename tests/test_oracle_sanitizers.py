"""The CPU oracle under AddressSanitizer + UndefinedBehaviorSanitizer: every entry point of
oracle/noc_oracle.c on small problems, both models, every optional output absent once, exact-size heap blocks.
compute-sanitizer is closed on the GPU pool, so this is the memory / UB tool that is left for the checker itself."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_is_clean_under_asan_and_ubsan():
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run(["sh", os.path.join(ROOT, "oracle", "san", "run.sh"), d], capture_output=True, text=True,
                           timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "ORACLE SANITIZE OK" in out, out[-3000:]
    assert "ERROR: AddressSanitizer" not in out and "runtime error" not in out, out[-3000:]
