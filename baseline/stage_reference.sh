#!/usr/bin/env bash
# Stage the reference's Python sources for the GPU box.  /root/reference does not exist there, but
# baseline/_ref/ (git-ignored, NOT gpurun-ignored) travels with the snapshot, so the tests that run
# the UNMODIFIED reference classes on backend "b200" against the real library
# (tests/test_gpu_plugin_reference.py) find them under $TASMANIA_REFERENCE = baseline/_ref.
# Nothing is installed and nothing enters the git history: a plain copy of src/tasmania (2 MB of
# Python; `pip install /root/reference` is impossible offline).
#
#   bash baseline/stage_reference.sh        (build container, before gpurun / at the end of a round)
set -eu
here="$(cd "$(dirname "$0")" && pwd)"
src="${TASMANIA_REFERENCE_SRC:-/root/reference/src}"
if [ ! -d "$src/tasmania" ]; then
  echo "stage_reference: $src/tasmania not found (nothing staged)" >&2
  exit 0
fi
rm -rf "$here/_ref/src"
mkdir -p "$here/_ref/src"
cp -r "$src/tasmania" "$here/_ref/src/tasmania"
find "$here/_ref" -name "__pycache__" -type d -prune -exec rm -rf {} +
echo "staged $(find "$here/_ref/src" -name '*.py' | wc -l) reference modules under $here/_ref/src"
