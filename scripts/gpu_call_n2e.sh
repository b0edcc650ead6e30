#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x -k "two or halo_exchange" ) > gpurun_out/n2e_pytest.log 2>&1; grep -E "passed|failed" gpurun_out/n2e_pytest.log | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu --skip-e2e ) > gpurun_out/n2e_bench_c3.json 2> gpurun_out/n2e_bench_c3.err; tail -c 700 gpurun_out/n2e_bench_c3.json; tail -2 gpurun_out/n2e_bench_c3.err
