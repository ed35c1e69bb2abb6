"""Every kernel family and every A/B form of a kernel (environment switches) once at ragged small sizes: the cases of
tools/sanitize_cases.py, run in a fresh process so that the switches cannot leak into other tests.  Parity is asserted
elsewhere; this keeps the seldom-used forms (loop-nest lane-pair kernel, one-thread banded kernel, 512-thread dense LU,
CTA kernels on warp-sized operators, ...) launching, converging and finite."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
def test_every_kernel_form_runs_on_ragged_sizes():
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "sanitize_cases.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "all cases ran" in r.stdout
    assert "FAILED" not in r.stdout
    assert r.stdout.count(" ok") >= 20, r.stdout
