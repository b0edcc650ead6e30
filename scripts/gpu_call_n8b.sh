#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time timeout 400 $TR --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --skip-e2e --skip-cpu ) > gpurun_out/n8b_bench_c3.json 2> gpurun_out/n8b_bench_c3.err; tail -c 900 gpurun_out/n8b_bench_c3.json; tail -3 gpurun_out/n8b_bench_c3.err
( time timeout 900 $TR --master-port 29514 bench.py --gpus 8 --config c4 --steps 3 --warmup 1 --skip-e2e --skip-cpu ) > gpurun_out/n8b_bench_c4.json 2> gpurun_out/n8b_bench_c4.err; tail -c 1500 gpurun_out/n8b_bench_c4.json; tail -3 gpurun_out/n8b_bench_c4.err
nvidia-smi --query-gpu=index,clocks.sm,power.draw,power.limit,temperature.gpu --format=csv
