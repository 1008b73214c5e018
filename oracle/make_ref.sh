#!/bin/bash
# Copies the UNMODIFIED reference script next to the oracle so that bench.py --impl reference can time the
# reference's own numpy/scipy path on the GPU box's host cores.  /root/reference does not
# exist on the GPU box; oracle/_ref/ is git-ignored but travels with gpurun, like the built .so files.
# The reference is a single Python script: there is nothing to compile.  Never edit the copy.
set -eu
here="$(cd "$(dirname "$0")" && pwd)"
src="${1:-/root/reference}/BalLeRMix+_v1.py"
[ -f "$src" ] || { echo "make_ref.sh: $src not found (nothing to do on a box without the reference)"; exit 0; }
mkdir -p "$here/_ref"
cp "$src" "$here/_ref/BalLeRMix+_v1.py"
( cd "$here/_ref" && sha256sum "BalLeRMix+_v1.py" > SHA256 )
echo "copied $src -> $here/_ref/"
