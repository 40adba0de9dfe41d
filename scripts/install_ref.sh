#!/usr/bin/env bash
# Copy the UNMODIFIED reference (a pure-Python script tree, no setup.py / pyproject: nothing to pip-install) into
# the git-ignored baseline/_ref/ so that it travels to the GPU box with the gpurun snapshot:
#   bench.py --impl reference        the reference's own CPU path on the host cores   (cpu_baseline.kind = reference)
#   bench.py (extras)                the reference run eagerly on the same B200       (eager_cuda_baseline)
#   scripts/run_reference_trainer.py the unchanged trainer driven through dropin/
# Nothing under baseline/_ref is product source; nothing in the package imports it.
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
DST="$HERE/baseline/_ref"
if [ ! -d "$SRC/models" ]; then
  echo "install_ref: no reference checkout at $SRC" >&2
  exit 1
fi
rm -rf "$DST"
mkdir -p "$DST"
for d in models regularization utils dataset train eval configs scripts; do
  cp -r "$SRC/$d" "$DST/$d"
done
cp "$SRC/config-defaults.yaml" "$SRC/LICENSE" "$SRC/README.md" "$DST/"
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo "unknown" ) > "$DST/.source_commit"
echo "install_ref: copied $(find "$DST" -type f | wc -l) files from $SRC to $DST"
