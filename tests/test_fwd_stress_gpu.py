"""Randomised forward configurations (incl. sequences long enough for the chunked block order, GQA, masks, ALiBi, softcap,
head dims up to 256): out and LSE against the oracle with the reference's 2x rule (tests/test_flash_attn.py:1121).
Long form: tools/stress_fwd.py (80 configurations clean on the B200 pool); this is a 25-configuration slice."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_randomised_forward_slice():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_fwd.py"), "21", "25"], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "stress done: 25/25 configurations clean" in r.stdout, r.stdout[-3000:]
