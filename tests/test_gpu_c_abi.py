"""The drop-in boundary called from plain C (no Python, no torch): tests/c/abi_smoke.c is compiled against
include/peritext_b200.h, linked to libperitext_b200.so and run on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_plain_c_caller(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
                           "-L", os.path.join(ROOT, "peritext_b200"), "-lperitext_b200", "-Wl,-rpath," + os.path.join(ROOT, "peritext_b200"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "aYXbc" in r.stdout and "converged" in r.stdout
