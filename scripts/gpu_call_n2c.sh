#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -k "two or halo_exchange" ) > gpurun_out/n2c_pytest.log 2>&1; grep -E "passed|failed" gpurun_out/n2c_pytest.log | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu --skip-e2e ) > gpurun_out/n2c_bench_c3.json 2> gpurun_out/n2c_bench_c3.err; tail -c 500 gpurun_out/n2c_bench_c3.json; tail -2 gpurun_out/n2c_bench_c3.err
( MVD_OVERLAP=0 timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu --skip-e2e ) > gpurun_out/n2c_bench_c3_noov.json 2> gpurun_out/n2c_bench_c3_noov.err; tail -c 500 gpurun_out/n2c_bench_c3_noov.json; tail -2 gpurun_out/n2c_bench_c3_noov.err
