#!/bin/bash
# Install the UNMODIFIED reference (abdo-eldesokey/RAFT-NCUP, /root/reference) into baseline/_ref for bench.py's
# `--impl reference` and `gpu_eager_baseline` legs.  baseline/_ref is git-ignored (never committed) but travels to the GPU box
# with the gpurun snapshot.  The reference ships no setup.py / pyproject (a directory of scripts), so the install goes through
# a copy under /tmp that adds a 10-line setup.py naming its `core` package; the reference's own files are not touched.
set -euo pipefail
REF=${RNC_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
[ -d "$REF/core" ] || { echo "no reference at $REF: nothing to install (the GPU box uses the prebuilt baseline/_ref)"; exit 0; }
TMP=$(mktemp -d /tmp/raft_ncup_ref.XXXXXX)
cp -r "$REF"/. "$TMP"/
cat > "$TMP/setup.py" <<'PY'
from setuptools import setup
setup(name="raft-ncup-reference", version="0+51ac387", description="unmodified abdo-eldesokey/RAFT-NCUP core/ (baseline arm)",
      packages=["core", "core.utils"], py_modules=[])
PY
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP" 2>&1 | tail -2
rm -rf "$TMP"
ls "$HERE/_ref/core" | head -20
