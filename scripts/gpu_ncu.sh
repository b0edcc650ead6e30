#!/bin/bash
# ncu --set full captures (with source-level stall sampling) of the pass kernels of the c3 plan (FFT length 540):
#   scripts/gpu_ncu.sh TAG [kernel-regex] [skip] [count]
# The kernel regex is matched against the demangled name; "Plan<\(int\)540" keeps the c3 tile kernels and drops the small plans of
# the PSF derivation.  skip: the spectra build (8 spectra x 3 launches) and the first iteration (4 views x 2 tiles x 9 passes).
mkdir -p gpurun_out
TAG=${1:-r02}
K=${2:-"Plan<\(int\)540"}
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${K}" -s ${3:-96} -c ${4:-9} \
    -o gpurun_out/ncu_${TAG} -f python scripts/prof_passes.py c3 1 > gpurun_out/ncu_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_${TAG}.log
ls -la gpurun_out/ncu_${TAG}.ncu-rep
