"""The example scripts (the reference's README tutorial and docs examples on the CUDA path) run to completion; each asserts its
own consistency checks (generated kernels == hand-written catalog kernels, finite fits)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("script", ["readme_tutorial.py", "multiple_time_series.py", "docs_examples.py"])
def test_example_runs(script):
    r = subprocess.run([sys.executable, os.path.join("examples", script)], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
