#!/bin/bash
# per-function SASS instruction counts of a built library. usage: tools/sass_sizes.sh [lib.so]
LIB=${1:-fasthevc_b200/libfasthevc_cuda.so}
LIB=$(readlink -f "$LIB")
T=$(mktemp -d); cd "$T"; cuobjdump -xelf all "$LIB" >/dev/null
nvdisasm -c *.cubin | awk '/^\s*\.text\./{name=$1} /^\s*\/\*[0-9a-f]+\*\//{cnt[name]++} END{for(n in cnt) print cnt[n], n}' | sort -rn | c++filt | head -40
rm -rf "$T"
