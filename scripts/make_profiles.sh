#!/bin/bash
# Runs on the GPU box: launch list of the bench command + full ncu captures of the pass kernels. Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu > gpurun_out/launches_bench.log 2>&1
# x kernels of the c3 tile (M = 540): skip spectra build (8) + warm-up (24), then P1, P5, P9
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:x_kernel.*int.540," -s 32 -c 3 \
    -o gpurun_out/prof_x_c3 -f python scripts/prof_passes.py c3 1 > gpurun_out/ncu_x.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:col_kernel.*int.540," -s 64 -c 3 \
    -o gpurun_out/prof_col_c3 -f python scripts/prof_passes.py c3 1 > gpurun_out/ncu_col.log 2>&1
python scripts/prof_passes.py c3 3 > gpurun_out/passes_c3.json 2>/dev/null
ls -la gpurun_out
