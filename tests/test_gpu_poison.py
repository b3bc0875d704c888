"""Uninitialised-memory guard: SRK_POISON=1 makes the library fill every fresh device buffer with NaN patterns, so a kernel that reads
memory nobody wrote (ragged last blocks, padding rows, stale work buffers) fails its parity test at once instead of depending on what
the allocator happens to hand out.  (compute-sanitizer is not available on the GPU pool.)  The parity suites are re-run in a
subprocess with the switch on."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("files,select", [
    (["tests/test_gpu_ekf_parity.py"], None),
    (["tests/test_gpu_ba_parity.py"], "small or sparse_scenes or skipped or pcg or ordered"),
])
def test_parity_suites_pass_on_poisoned_buffers(files, select):
    env = dict(os.environ, SRK_POISON="1")
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider"] + files + (["-k", select] if select else [])
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
