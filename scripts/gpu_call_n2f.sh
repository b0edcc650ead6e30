#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x -k "two or halo_exchange or loop_parity or golden" ) > gpurun_out/n2f_pytest.log 2>&1; grep -E "passed|failed" gpurun_out/n2f_pytest.log | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for sp in 0 1; do
( MVD_SPLIT_P1=$sp timeout 300 $TR --master-port 2951$sp bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu --skip-e2e ) > gpurun_out/n2f_sp$sp.json 2> gpurun_out/n2f_sp$sp.err; tail -c 500 gpurun_out/n2f_sp$sp.json | head -c 420; echo; tail -1 gpurun_out/n2f_sp$sp.err | cut -c1-200
done
