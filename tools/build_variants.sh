#!/bin/bash
# Build tuning variants of libcodeps_photo.so: name:TILE_Y:STRIP:MIN_CTAS
set -e
cd "$(dirname "$0")/.."
for v in "$@"; do
  IFS=: read name ty strip ctas <<< "$v"
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -cudart static \
    -I include -I codeps_b200/csrc -DCDP_TILE_Y=$ty -DCDP_STRIP=$strip -DCDP_PHOTO_MIN_CTAS=$ctas \
    -o codeps_b200/variants/lib_$name.so codeps_b200/csrc/cdp_api.cu
  echo "built $name"
done
