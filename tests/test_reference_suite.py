"""The reference's OWN test-suite, run in place against this package.

Only where the reference tree is mounted (the build container); it is never copied.  A pytest plug-in
(tests/ref_alias_plugin.py) makes ``import turtlemd...`` resolve to ``turtlemd_b200...`` in a child
process, which then collects /root/reference/tests unchanged.  Without a GPU the tests that reach a kernel
must fail with NoDeviceError -- there is no CPU fallback -- and every other test must pass: that is the
host-side API contract (constructors, attributes, exceptions, containers, stop conditions, I/O helpers).
With a GPU all of them must pass.
"""
from __future__ import annotations

import os
import pathlib
import re
import subprocess
import sys

import pytest

from conftest import REPO

REF_TESTS = pathlib.Path("/root/reference/tests")


@pytest.mark.skipif(not REF_TESTS.is_dir(), reason="the reference tree is not mounted here")
def test_reference_suite_against_this_package(tmp_path):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(REPO / "tests"), str(REPO)]))
    cmd = [sys.executable, "-m", "pytest", "-p", "ref_alias_plugin", "-p", "no:cacheprovider", "--rootdir", str(tmp_path),
           "-q", "-rf", str(REF_TESTS)]
    out = subprocess.run(cmd, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=900).stdout
    summary = out.strip().splitlines()[-1]
    passed = int(re.search(r"(\d+) passed", summary).group(1))
    failed = int(m.group(1)) if (m := re.search(r"(\d+) failed", summary)) else 0
    assert passed + failed == 63, summary  # the reference ships 63 tests
    from turtlemd_b200 import _lib

    has_gpu = False
    try:
        has_gpu = _lib.device_count() > 0
    except Exception:  # noqa: BLE001
        pass
    if has_gpu:
        assert failed == 0, out[-3000:]
    else:
        # every failure is the loud "no device" error of a test that reaches a kernel
        blocks = re.split(r"\n_{5,} (?=test)", out)
        loud = sum(1 for b in blocks if "NoDeviceError" in b and re.match(r"test\w* _{5,}", b))
        assert failed == loud, out[-3000:]
        assert passed >= 47, summary
