#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/n4_bench_c3.json 2> gpurun_out/n4_bench_c3.err; tail -c 700 gpurun_out/n4_bench_c3.json; tail -1 gpurun_out/n4_bench_c3.err | cut -c1-200
