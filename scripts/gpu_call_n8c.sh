#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for ov in 1 0; do
( MVD_OVERLAP=$ov timeout 300 $TR --master-port 2951$ov bench.py --gpus 8 --steps 20 --warmup 3 --skip-cpu --skip-e2e --skip-parity ) > gpurun_out/n8c_ov$ov.json 2> gpurun_out/n8c_ov$ov.err; tail -c 600 gpurun_out/n8c_ov$ov.json; tail -1 gpurun_out/n8c_ov$ov.err
done
