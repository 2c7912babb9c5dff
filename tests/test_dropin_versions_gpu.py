"""Drop-in on BOTH versions of the reference's FA2 interface (installed 2.8.3; the reference tree's 2.8.4 from
baseline/_ref, SURVEY 7 'Cross-version ABI') and on the reference's FA3 interface (hopper/flash_attn_interface.py over
TORCH_LIBRARY(flash_attn_3)).  Each file is loaded unmodified in a fresh interpreter (tests/dropin_scripts/run_ref_interface.py):
forward + backward, varlen, kv-cache append, torch.compile(fullgraph=True) through the reference's own custom ops, fp8
de-scales / paged cache / combine / scheduler metadata through the FA3 names."""
import importlib.util
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "tests", "dropin_scripts", "run_ref_interface.py")


def _run(kind, path):
    if not os.path.exists(path):
        pytest.skip(f"{path} is not present on this box")
    r = subprocess.run([sys.executable, SCRIPT, kind, path], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "OK " in r.stdout, (r.stdout[-2000:] + "\n" + r.stderr[-4000:])


def test_installed_2_8_3_interface():
    spec = importlib.util.find_spec("flash_attn")
    if spec is None:
        pytest.skip("flash_attn is not installed in this image")
    _run("fa2", os.path.join(list(spec.submodule_search_locations)[0], "flash_attn_interface.py"))


def test_reference_tree_2_8_4_interface():
    _run("fa2", os.path.join(ROOT, "baseline", "_ref", "flash_attn", "flash_attn_interface.py"))


def test_reference_fa3_interface():
    _run("fa3", os.path.join(ROOT, "baseline", "_ref", "hopper", "flash_attn_interface.py"))
