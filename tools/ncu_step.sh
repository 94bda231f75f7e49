#!/bin/bash
# Full ncu capture of one step's own kernels (third step of tools/prof_step.py):
#   tools/ncu_step.sh NAME   ->  gpurun_out/NAME.ncu-rep   (run on the GPU box via gpurun)
# Per step the library launches 7 kernels (pyramid incl. intrinsics table, photo, finalize, smooth main / finalize,
# smooth bwd, depth grad); two warm-up steps are skipped.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:cdp_ -s 14 -c 7 -f \
    -o "gpurun_out/${1:-prof_step}" python tools/prof_step.py
