#!/bin/bash
# Installs the UNMODIFIED reference (lueckem/SPoNet, /root/reference) into baseline/_ref (git-ignored, travels
# to the GPU box with the snapshot) so that `bench.py --impl reference` can time its own numba +
# multiprocessing path there.  Build container only (/root/reference does not exist on the GPU box).
#
# The reference builds with poetry-core, which is neither installed nor in the offline wheelhouse; its
# pyproject.toml carries a complete PEP 621 [project] table, so a copy under /tmp gets the build backend
# swapped for setuptools (packaging metadata only -- no file of the `sponet` package is touched) and is
# installed from there.  Dependencies (numba, networkx, scipy, tqdm) are already in the image: --no-deps.
set -e
cd "$(dirname "$0")/.."
REF=${1:-/root/reference}
[ -d "$REF/sponet" ] || { echo "no reference at $REF"; exit 1; }
TMP=$(mktemp -d /tmp/sponet_ref.XXXXXX)
cp -r "$REF"/. "$TMP"/
python - "$TMP/pyproject.toml" <<'PY'
import re, sys
p = sys.argv[1]
s = open(p).read()
s = re.sub(r'\[build-system\].*?(?=\n\[|\Z)', '[build-system]\nrequires = ["setuptools"]\nbuild-backend = "setuptools.build_meta"\n', s, flags=re.S)
s = re.sub(r'\n\[tool\.poetry[^\n]*\].*?(?=\n\[(?!tool\.poetry)|\Z)', '\n', s, flags=re.S)
s = s.replace('"tqdm (>=4.67.1,<5.0.0)"', '"tqdm>=4.67.1,<5.0.0"')
s = s.replace('license = "GPL-3.0-or-later"', 'license = {text = "GPL-3.0-or-later"}')
s += '\n[tool.setuptools.packages.find]\ninclude = ["sponet*"]\n'
open(p, "w").write(s)
PY
rm -rf baseline/_ref
mkdir -p baseline
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target baseline/_ref "$TMP" 2>&1 | tail -3
rm -rf "$TMP"
# the installed package must be byte-identical to the reference's
diff -r -q "$REF/sponet" baseline/_ref/sponet -x __pycache__ && echo "baseline/_ref/sponet == $REF/sponet"
