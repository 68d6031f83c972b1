#!/usr/bin/env bash
# Copies the reference's own sources for the learner hot path into the git-ignored baseline/_ref/ so that the GPU box
# (which receives /root/repo only, never /root/reference) can time the UNMODIFIED reference as the CPU arm
# (bench.py --impl reference, cpu_baseline.kind = "reference").
#
#   tools/make_baseline_ref.sh [/path/to/reference]        (default: $ICRL_REFERENCE_ROOT or /root/reference)
#
# What is copied (sources only, byte for byte, no edits): the vendored stable_baselines3 package (its __init__ imports every
# algorithm, so the package travels whole) and icrl/constraint_net.py.  They are imported through oracle/ref_shim.py (stub
# gym / matplotlib modules).  baseline/_ref/ is listed in .gitignore -- it never enters the history -- but not in
# .gpurunignore, so it travels to the GPU box like the built .so files.
set -euo pipefail
SRC="${1:-${ICRL_REFERENCE_ROOT:-/root/reference}}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
DST="$HERE/baseline/_ref"
if [ ! -d "$SRC/stable_baselines3" ] || [ ! -f "$SRC/icrl/constraint_net.py" ]; then
    echo "make_baseline_ref: no reference checkout at $SRC" >&2
    exit 1
fi
rm -rf "$DST"
mkdir -p "$DST/icrl"
cp -r "$SRC/stable_baselines3" "$DST/stable_baselines3"
cp "$SRC/icrl/constraint_net.py" "$DST/icrl/constraint_net.py"
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$SRC" && find stable_baselines3 icrl/constraint_net.py -name '*.py' -type f | sort | xargs sha256sum ) > "$DST/SOURCES.sha256"
echo "baseline/_ref: $(find "$DST" -name '*.py' | wc -l) files from $SRC"
