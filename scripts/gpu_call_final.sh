#!/bin/bash
# round-2 final state: bench lines (c3 default, c2), GPU tests, smoke, ncu launch list of the bench command, ncu --set full of the nine passes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/final_smi.txt 2>&1
( time timeout 600 python bench.py ) > gpurun_out/final_bench_c3.json 2> gpurun_out/final_bench_c3.err; tail -c 1200 gpurun_out/final_bench_c3.json; tail -2 gpurun_out/final_bench_c3.err
( time timeout 300 python bench.py --config c2 --steps 5 --skip-cpu ) > gpurun_out/final_bench_c2.json 2> gpurun_out/final_bench_c2.err; tail -c 500 gpurun_out/final_bench_c2.json; tail -2 gpurun_out/final_bench_c2.err
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/final_pytest.log 2>&1; grep -E "passed|failed" gpurun_out/final_pytest.log | tail -2
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --skip-parity > gpurun_out/final_launches.log 2>&1; tail -2 gpurun_out/final_launches.log | cut -c1-200
timeout 500 bash scripts/gpu_ncu.sh final
