"""Runs the tests that need the JavaScript engine (tests/jsref_live_cases.py: the oracle, the suite generators and the NPY format
against the live reference; tests/js_shim_cases.py: the JS shim executed beside the reference's modules) in a CHILD pytest process.

QJSEngine is loaded through ctypes; it is somebody else's JIT-ing interpreter driven through a hand-written ABI, and one of its bugs
was a segmentation fault (oracle/jsref/qjs.py).  A crash there must not take this test session down with it: a child that dies from a
signal is reported as a skip with the signal number, a child that fails is a failure with its output.
"""
import os
import subprocess
import sys

import pytest

from oracle.jsref import qjs

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not qjs.available(), reason="needs /root/reference and the Qt 6 QJSEngine shipped with Nsight Compute")


@pytest.mark.parametrize("cases", ["jsref_live_cases.py", "js_shim_cases.py"])
def test_engine_backed_cases_in_a_child_process(cases):
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, cases), "-q", "-x", "-p", "no:cacheprovider"],
                       capture_output=True, text=True, timeout=900, cwd=os.path.dirname(HERE))
    if r.returncode < 0:
        pytest.skip("the JavaScript engine crashed (signal %d); tail: %s" % (-r.returncode, (r.stdout + r.stderr)[-400:]))
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]
