#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 1500 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --skip-parity > gpurun_out/final_launches.log 2>&1; tail -2 gpurun_out/final_launches.log | cut -c1-300; wc -l gpurun_out/final_launches.csv
