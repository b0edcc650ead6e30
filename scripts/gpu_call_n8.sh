#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time timeout 600 $TR --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/n8_bench_c3.json 2> gpurun_out/n8_bench_c3.err; tail -c 1800 gpurun_out/n8_bench_c3.json; tail -3 gpurun_out/n8_bench_c3.err
( time BENCH_GRID=8x1 timeout 400 $TR --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --skip-e2e --skip-cpu ) > gpurun_out/n8_bench_c3_8x1.json 2> gpurun_out/n8_bench_c3_8x1.err; tail -c 500 gpurun_out/n8_bench_c3_8x1.json; tail -3 gpurun_out/n8_bench_c3_8x1.err
( time timeout 600 $TR --master-port 29513 bench.py --gpus 8 --config c5 --steps 3 --warmup 1 --skip-e2e --skip-cpu ) > gpurun_out/n8_bench_c5.json 2> gpurun_out/n8_bench_c5.err; tail -c 1200 gpurun_out/n8_bench_c5.json; tail -3 gpurun_out/n8_bench_c5.err
( time timeout 900 $TR --master-port 29514 bench.py --gpus 8 --config c4 --steps 3 --warmup 1 --skip-e2e --skip-cpu ) > gpurun_out/n8_bench_c4.json 2> gpurun_out/n8_bench_c4.err; tail -c 1200 gpurun_out/n8_bench_c4.json; tail -3 gpurun_out/n8_bench_c4.err
