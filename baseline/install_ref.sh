#!/bin/bash
# Populates baseline/_ref (git-ignored, travels to the GPU box) with UNMODIFIED reference files -- nothing under
# baseline/_ref is ever edited:
#   * the reference's Blackwell package `flash_attn.cute` (FA4), pip-installed from a copy of its source directory
#     (the reference tree is read-only and setuptools writes egg-info next to pyproject.toml) -- the same-box comparator
#     of tools/baseline_probe.py (with the DSL binding shims of tools/fa4_shim.py);
#   * the reference tree's 2.8.4 `flash_attn/flash_attn_interface.py` and `hopper/flash_attn_interface.py`, which
#     tests/test_dropin_versions_gpu.py loads on top of this package's operator module;
#   * (optional, slow) CUTLASS example 77: baseline/build_ex77.sh.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${REF:-/root/reference}
rm -rf /tmp/fa4src && cp -r "$REF/flash_attn/cute" /tmp/fa4src
SETUPTOOLS_SCM_PRETEND_VERSION=4.0.0 python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
  --target "$HERE/_ref" /tmp/fa4src
cp "$REF/flash_attn/flash_attn_interface.py" "$HERE/_ref/flash_attn/flash_attn_interface.py"
mkdir -p "$HERE/_ref/hopper" && cp "$REF/hopper/flash_attn_interface.py" "$HERE/_ref/hopper/flash_attn_interface.py"
ls "$HERE/_ref" "$HERE/_ref/flash_attn" "$HERE/_ref/hopper"
