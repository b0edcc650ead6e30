#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for gap in 0 200 1000 3000; do
BENCH_EXCHANGE_ONLY=1 BENCH_EXCHANGE_GAP_US=$gap timeout 300 $TR --master-port 2951$((gap/1000)) bench.py --gpus 2 --skip-cpu --skip-e2e 2>/dev/null | tail -1
done | tee gpurun_out/exchange_gap.jsonl
nvidia-smi nvlink -s -i 0 2>&1 | head -8
nvidia-smi -q -i 0 2>/dev/null | grep -i -A6 "nvlink\|link power" | head -30
