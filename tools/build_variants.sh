#!/bin/bash
# Build tuning variants of libcodeps_photo.so:  name:"-DFLAG=.. -DFLAG=.."
set -e
cd "$(dirname "$0")/.."
mkdir -p codeps_b200/variants
for v in "$@"; do
  name="${v%%:*}"; flags="${v#*:}"
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -cudart static \
    -I include -I codeps_b200/csrc $flags -Xptxas -v -o codeps_b200/variants/lib_$name.so codeps_b200/csrc/cdp_api.cu 2>&1 \
    | grep -A2 "cdp_photo_kernelILb1" | grep -E "spill" | sed "s/^/$name: /"
done
