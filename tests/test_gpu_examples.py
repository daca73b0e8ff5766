"""The example scripts run end to end (small sizes) and their self-checks pass."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("script,args", [
    ("locate_earthquakes.py", ["--chains", "256", "--proposals", "600"]),
    ("linear_tomography.py", ["--grid", "16", "--rays", "600", "--chains", "128", "--proposals", "200"]),
    ("parallel_tempering.py", ["--chains", "16", "--proposals", "200"]),
])
def test_example_runs(script, args):
    env = {k: v for k, v in os.environ.items() if k not in ("LOCAL_RANK", "RANK", "WORLD_SIZE")}   # single process
    res = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script), *args], env=env,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
