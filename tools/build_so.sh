#!/bin/bash
# Builds libfasthevc_cuda.so (and with "prof" the FHC_PROFILE variant) and prints register use + SASS instruction count.
# usage: tools/build_so.sh [prof] [extra nvcc flags...]
set -e
cd "$(dirname "$0")/.."
OUT=fasthevc_b200/libfasthevc_cuda.so
EXTRA=()
if [ "$1" = "prof" ]; then OUT=fasthevc_b200/libfasthevc_cuda_prof.so; EXTRA+=(-DFHC_PROFILE); shift; fi
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v \
  "${EXTRA[@]}" "$@" -o "$OUT" fasthevc_b200/csrc/fasthevc_cuda.cu 2>&1 | grep -E "error|Used [0-9]+ registers|warning.*stack" || true
T=$(mktemp -d)
(cd "$T" && cuobjdump -xelf all "$OLDPWD/$OUT" >/dev/null && echo "SASS instructions: $(nvdisasm -c *.cubin | grep -c '^\s*/\*[0-9a-f]*\*/')")
rm -rf "$T"
