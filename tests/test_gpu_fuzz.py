"""Randomised differential tests (tools/fuzz_parity.py, tools/fuzz_ukf.py) as part of the suite: fixed seeds, shipped shapes only
(UDE_JIT=0 turns a shape without a pre-compiled kernel into a counted refusal instead of an nvcc run)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args):
    env = dict(os.environ, UDE_JIT="0")
    r = subprocess.run([sys.executable] + args, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    return r.stdout


def test_random_problems_against_the_oracle():
    out = _run(["tools/fuzz_parity.py", "--n", "300", "--seed", "7", "--forward", "--sparse", "--adam"])
    assert "failures 0" in out


def test_random_problems_fp32_mode():
    out = _run(["tools/fuzz_parity.py", "--n", "150", "--seed", "8", "--dtype", "f32"])
    assert "failures 0" in out


def test_random_filters_against_the_twin():
    out = _run(["tools/fuzz_ukf.py", "--n", "30", "--seed", "9"])
    assert "failures 0" in out
