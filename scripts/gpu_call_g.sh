#!/bin/bash
mkdir -p gpurun_out
( timeout 400 bash scripts/sweep_env.sh MVD_XWARP "2 1 0 2" c3 2 ) 2>&1 | tee gpurun_out/ab_g.txt
timeout 300 python bench.py --skip-e2e --skip-cpu --steps 5 --warmup 3 > gpurun_out/g_bench_c3.json 2> gpurun_out/g_bench_c3.err; tail -c 700 gpurun_out/g_bench_c3.json; tail -3 gpurun_out/g_bench_c3.err
timeout 300 python bench.py --config c2 --skip-e2e --skip-cpu --steps 5 --warmup 3 > gpurun_out/g_bench_c2.json 2> gpurun_out/g_bench_c2.err; tail -c 300 gpurun_out/g_bench_c2.json; tail -3 gpurun_out/g_bench_c2.err
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/g_pytest.log 2>&1; tail -5 gpurun_out/g_pytest.log
