#!/bin/bash
# GPU call 1 (round 2): validation of the new parity tests / bench parity leg + micro-benchmarks
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py ) > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err
tail -c 3000 gpurun_out/bench_c3_n1.json; tail -5 gpurun_out/bench_c3_n1.err
( time timeout 600 python bench.py --config c2 --steps 5 --skip-cpu ) > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err
tail -c 1500 gpurun_out/bench_c2_n1.json; tail -5 gpurun_out/bench_c2_n1.err
timeout 120 scripts/microbench/mb > gpurun_out/microbench.jsonl 2>&1
cat gpurun_out/microbench.jsonl
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
