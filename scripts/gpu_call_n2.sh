#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/n2_smi.txt
( time timeout 600 python -m pytest tests -m gpu -q -k "two or halo_exchange" ) > gpurun_out/n2_pytest.log 2>&1; tail -5 gpurun_out/n2_pytest.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu ) > gpurun_out/n2_bench_c3.json 2> gpurun_out/n2_bench_c3.err; tail -c 2500 gpurun_out/n2_bench_c3.json; tail -3 gpurun_out/n2_bench_c3.err
