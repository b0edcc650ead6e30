#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time timeout 400 $TR --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/n8d_bench_c3.json 2> gpurun_out/n8d_bench_c3.err; tail -c 1500 gpurun_out/n8d_bench_c3.json; tail -2 gpurun_out/n8d_bench_c3.err
