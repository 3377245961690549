#!/usr/bin/env bash
# Builds the UNMODIFIED reference (SonSang/dmesh2_renderer) CUDA extension for sm_100a into
# baseline/_ref/ (git-ignored, travels to the GPU box with gpurun).  The reference tree is
# read-only and `setup.py` includes an un-vendored, never-used glm header, so we build from a
# scratch copy under /tmp with an empty glm stub.  No reference source is
# copied into the repository history: baseline/_ref/ is in .gitignore.
set -euo pipefail
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then echo "reference tree $REF not present; nothing to build"; exit 0; fi
TMP=$(mktemp -d /tmp/dmesh2_ref_build.XXXXXX)
cp -r "$REF"/. "$TMP"/ && chmod -R u+w "$TMP"
mkdir -p "$TMP/third_party/glm/glm" && echo '#pragma once' > "$TMP/third_party/glm/glm/glm.hpp"
rm -rf "$OUT" && mkdir -p "$OUT"
cd "$TMP"
MAX_JOBS=${MAX_JOBS:-6} TORCH_CUDA_ARCH_LIST=10.0a python -m pip install --no-index --no-build-isolation \
    --no-deps --find-links /opt/wheelhouse --target "$OUT" "$TMP" 2>&1 | tail -n 15
ls -la "$OUT" "$OUT/dmesh2_renderer" || true
rm -rf "$TMP"
