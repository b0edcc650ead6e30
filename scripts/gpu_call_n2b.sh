#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -k "two or halo_exchange" ) > gpurun_out/n2b_pytest.log 2>&1; tail -4 gpurun_out/n2b_pytest.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu --skip-e2e ) > gpurun_out/n2b_bench_c3.json 2> gpurun_out/n2b_bench_c3.err; tail -c 600 gpurun_out/n2b_bench_c3.json; tail -3 gpurun_out/n2b_bench_c3.err
BENCH_EXCHANGE_ONLY=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --skip-cpu --skip-e2e 2>/dev/null | tail -1
