#!/bin/bash
# round-2 state check: bench c3 (full line), bench c2, GPU tests, ncu full capture of the nine c3 passes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 600 python bench.py ) > gpurun_out/a_bench_c3.json 2> gpurun_out/a_bench_c3.err
tail -c 2500 gpurun_out/a_bench_c3.json; tail -3 gpurun_out/a_bench_c3.err
( time timeout 300 python bench.py --config c2 --steps 5 --skip-cpu ) > gpurun_out/a_bench_c2.json 2> gpurun_out/a_bench_c2.err
tail -c 1200 gpurun_out/a_bench_c2.json; tail -3 gpurun_out/a_bench_c2.err
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/a_pytest.log 2>&1; tail -5 gpurun_out/a_pytest.log
timeout 600 bash scripts/gpu_ncu.sh a
