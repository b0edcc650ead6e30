#!/bin/bash
mkdir -p gpurun_out
( timeout 300 bash scripts/sweep_env.sh MVD_PF_Y "296 444 592" c3 2
  timeout 300 bash scripts/sweep_env.sh MVD_PF_X "19 37" c3 2 ) 2>&1 | tee gpurun_out/ab_e.txt
timeout 300 python bench.py --config c2 --skip-e2e --skip-cpu --steps 5 --warmup 3 > gpurun_out/e_bench_c2.json 2> gpurun_out/e_bench_c2.err; tail -c 300 gpurun_out/e_bench_c2.json; tail -3 gpurun_out/e_bench_c2.err
timeout 600 bash scripts/gpu_ncu.sh e
