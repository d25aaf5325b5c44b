#!/bin/sh
# Populates baseline/_ref/ with the UNMODIFIED reference package so that `bench.py --impl reference` and the oracle shim
# (oracle/ref_shim.py: $BPMULT_REF, /root/reference, baseline/_ref in that order) can run the reference's own modules on a GPU box,
# where /root/reference does not exist.  baseline/_ref is git-ignored (the reference sources never enter this repository's history)
# but is NOT gpurun-ignored, so it travels with the working tree.
#
# The reference is pure Python without setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` has nothing
# to build ("neither 'setup.py' nor 'pyproject.toml' found"); the package directory is copied as it is instead.
set -e
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")/.." && pwd)"
DST="$HERE/baseline/_ref"
if [ ! -d "$SRC/bpmult/models" ]; then
  echo "install_reference: no reference tree at $SRC" >&2
  exit 1
fi
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$SRC/bpmult" "$DST/bpmult"
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
[ -f "$SRC/README.md" ] && cp "$SRC/README.md" "$DST/README.md"
( cd "$SRC" && find bpmult -name '*.py' | sort | xargs sha256sum ) > "$DST/SHA256SUMS"
echo "install_reference: $(find "$DST/bpmult" -name '*.py' | wc -l) files -> $DST"
