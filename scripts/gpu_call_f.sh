#!/bin/bash
mkdir -p gpurun_out
( timeout 300 bash scripts/sweep_env.sh MVD_XWARP "1 0 1" c3 2 ) 2>&1 | tee gpurun_out/ab_f.txt
timeout 300 python bench.py --skip-e2e --skip-cpu --steps 5 --warmup 3 > gpurun_out/f_bench_c3.json 2> gpurun_out/f_bench_c3.err; tail -c 700 gpurun_out/f_bench_c3.json; tail -3 gpurun_out/f_bench_c3.err
timeout 300 python bench.py --config c2 --skip-e2e --skip-cpu --steps 5 --warmup 3 > gpurun_out/f_bench_c2.json 2> gpurun_out/f_bench_c2.err; tail -c 300 gpurun_out/f_bench_c2.json; tail -3 gpurun_out/f_bench_c2.err
