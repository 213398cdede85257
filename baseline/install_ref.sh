#!/bin/sh
# Install the UNMODIFIED reference package (irkri/fruits 1.0.0) into baseline/_ref so that
# `bench.py --impl reference` can time the real numba implementation on the GPU box.
#
# The reference's pyproject.toml names poetry-core as its build backend, which is not in
# the image (and there is no network), so the stock `pip install /root/reference` fails
# with "No module named 'poetry'".  This script installs from a copy under /tmp whose
# *build metadata only* is replaced by an equivalent setuptools description; every file
# of the `fruits` package itself is the reference's, byte for byte (checked below).
# baseline/_ref is git-ignored (never part of the history) but travels with gpurun.
set -e
REF=${FRUITS_REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
TMP=$(mktemp -d /tmp/fruits_ref.XXXXXX)
cp -r "$REF/fruits" "$TMP/fruits"
cat > "$TMP/pyproject.toml" <<'TOML'
[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"
[project]
name = "fruits"
version = "1.0.0"
description = "Feature Extraction Using Iterated Sums (reference, unmodified sources)"
[tool.setuptools.packages.find]
include = ["fruits*"]
TOML
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP" >/dev/null
diff -r -x __pycache__ "$REF/fruits" "$HERE/_ref/fruits" && echo "baseline/_ref/fruits is identical to $REF/fruits"
rm -rf "$TMP"
