#!/bin/bash
# ncu --set full captures (with source-level stall sampling) of the pass kernels of the c3 plan
mkdir -p gpurun_out
TAG=${1:-r02}
K=${2:-"x_kernel|col_kernel"}
# skip the spectra build and the first (warm-up) iteration: 8 spectra x 3 launches + 1 iteration x 4 views x 2 tiles x 9 passes
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${K}" -s ${3:-100} -c ${4:-9} \
    -o gpurun_out/ncu_${TAG} -f python scripts/prof_passes.py c3 1 > gpurun_out/ncu_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_${TAG}.log
ls -la gpurun_out/ncu_${TAG}.ncu-rep
