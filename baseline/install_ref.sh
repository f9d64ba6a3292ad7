#!/usr/bin/env bash
# Installs the UNMODIFIED reference (zgcr/SimpleAICV_pytorch_training_examples) into baseline/_ref
# (git-ignored; travels to the GPU box with gpurun).  The reference ships no setup.py / pyproject,
# and /root/reference is read-only, so the install runs from a copy under /tmp with a minimal
# setup.py that only lists its two import roots (SimpleAICV, tools).  No source file is edited.
set -euo pipefail
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
TMP=$(mktemp -d /tmp/refcopy.XXXXXX)
cp -r "$REF/SimpleAICV" "$REF/tools" "$TMP/"
cat > "$TMP/setup.py" <<'PY'
from setuptools import setup, find_namespace_packages
setup(name='simpleaicv_reference', version='0.0.0',
      packages=find_namespace_packages(include=['SimpleAICV', 'SimpleAICV.*', 'tools', 'tools.*']))
PY
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP" 2>&1 | tail -3
rm -rf "$TMP"
python - <<PY
import sys; sys.path.insert(0, "$HERE/_ref")
from SimpleAICV.classification import backbones
print('reference installed:', len([k for k in backbones.__dict__ if not k.startswith('_')]), 'backbone symbols')
PY
