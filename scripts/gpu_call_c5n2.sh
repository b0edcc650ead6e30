#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config c5 --steps 2 --warmup 1 --skip-cpu --skip-e2e ) > gpurun_out/c5n2.json 2> gpurun_out/c5n2.err; tail -c 1500 gpurun_out/c5n2.json; tail -5 gpurun_out/c5n2.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
free -g | head -2
