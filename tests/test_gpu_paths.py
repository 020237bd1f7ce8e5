"""Every kernel family on small cases, cross-checked against each other (tools/sanitize_smoke.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_all_paths_agree_on_small_cases():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_smoke.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "sanitize_smoke ok" in r.stdout, r.stdout + r.stderr
