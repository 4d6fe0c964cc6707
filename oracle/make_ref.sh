#!/usr/bin/env bash
# TEST / BASELINE INFRASTRUCTURE ONLY.
# Stages the UNMODIFIED reference (XiYe20/VPTR, pure Python) into the git-ignored oracle/_ref/ so that it travels to the
# GPU box with the gpurun snapshot (/root/reference does not exist there).  Nothing is patched: the files are byte-for-byte
# copies; the two adapters the reference needs on this image (a 2-symbol timm shim and CPU-device positional embeddings)
# are applied at import time by oracle/ref_loader.py.  Used by
#   * bench.py --impl reference and bench.py's cpu_baseline leg (kind "reference": train_NAR.single_iter / train_FAR.single_iter),
#   * tests/test_gpu_dropin.py (the reference's own single_iter driving the vptr_b200 modules on the B200),
#   * tests/golden/make_golden.py (fixture generation).
# Never imported by the product package vptr_b200/.
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DST="$HERE/_ref"
if [ ! -d "$SRC/model" ]; then
  echo "make_ref.sh: reference tree not found at $SRC (prebuilt oracle/_ref is used as is)" >&2
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/model" "$DST/utils"
cp "$SRC"/model/*.py "$DST/model/"
cp "$SRC"/utils/*.py "$DST/utils/"
cp "$SRC"/train_*.py "$DST/"
[ -f "$SRC/LICENSE" ] && cp "$SRC/LICENSE" "$DST/"
( cd "$SRC" && find model utils -name '*.py' -print0 | sort -z | xargs -0 sha256sum; sha256sum train_*.py ) > "$DST/SHA256SUMS"
echo "staged $(find "$DST" -name '*.py' | wc -l) reference files into $DST"
