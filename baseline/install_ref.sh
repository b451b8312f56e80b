#!/bin/bash
# Installs the UNMODIFIED reference (openmcmc 1.0.7, pure Python) into baseline/_ref (git-ignored; travels to the GPU
# box with the gpurun snapshot) for `bench.py --impl reference` and tests/golden/make_golden.py.
#
# `pip install --target baseline/_ref /root/reference` fails in this image: the reference's build backend is
# poetry-core, which is neither installed nor in /opt/wheelhouse (ModuleNotFoundError: No module named 'poetry').
# The package sources are installed as they are from a copy under /tmp whose pyproject.toml names setuptools as the
# build backend instead (build metadata only; `diff -r baseline/_ref/openmcmc /root/reference/src/openmcmc` is empty).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
[ -d "$REF/src/openmcmc" ] || { echo "no reference at $REF"; exit 0; }
TMP="$(mktemp -d)"
cp -r "$REF" "$TMP/ref"
python - "$TMP/ref/pyproject.toml" <<'EOF'
import sys
path = sys.argv[1]
s = open(path).read()
s = s.replace('requires = ["poetry-core>=1.0.0"]', 'requires = ["setuptools"]')
s = s.replace('build-backend = "poetry.core.masonry.api"', 'build-backend = "setuptools.build_meta"')
s += '\n[project]\nname = "openmcmc"\nversion = "1.0.7"\n\n[tool.setuptools.packages.find]\nwhere = ["src"]\n'
open(path, "w").write(s)
EOF
rm -rf "$ROOT/baseline/_ref"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" "$TMP/ref"
rm -rf "$TMP"
diff -r -x __pycache__ "$ROOT/baseline/_ref/openmcmc" "$REF/src/openmcmc" && echo "baseline/_ref == reference sources"
