#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python bench.py ) > gpurun_out/last_bench_c3.json 2> gpurun_out/last_bench_c3.err; tail -c 1500 gpurun_out/last_bench_c3.json; tail -2 gpurun_out/last_bench_c3.err
( timeout 200 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/last_ref.json 2> gpurun_out/last_ref.err; tail -c 400 gpurun_out/last_ref.json
