"""The C-ABI driven from plain C (tests/c_harness/fdx_harness.c): no torch, no Python in the loop - what an
XLA FFI handler or any other non-Python host does.  build() compiles the harness."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_plain_c_harness_drives_conv3x3_through_the_c_abi():
    exe = os.path.join(HERE, "c_harness", "fdx_harness")
    if not os.path.exists(exe):
        r = subprocess.run(["make", "-C", os.path.join(HERE, "c_harness")], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "HARNESS OK" in r.stdout, r.stdout + r.stderr
    assert "rejected bad input with status -" in r.stdout
