#!/bin/bash
mkdir -p gpurun_out
L=multiview-reconstruction_b200/libmvdecon.so
( bash scripts/ab_libs.sh c3 2 $L variants/libmvdecon_w8.so variants/libmvdecon_xlp2.so variants/libmvdecon_xlp8.so variants/libmvdecon_c3018.so
  bash scripts/sweep_env.sh MVD_PF_Z "37 74 296 592" c3 2
  bash scripts/sweep_env.sh MVD_PF_X "0 37 148 296" c3 2
  bash scripts/sweep_env.sh MVD_CHUNK_MB "40 80" c3 2 ) 2>&1 | tee gpurun_out/ab_c.txt
python bench.py --config c2 --skip-e2e --skip-cpu --steps 5 --warmup 3 > gpurun_out/c_bench_c2.json 2> gpurun_out/c_bench_c2.err; tail -c 900 gpurun_out/c_bench_c2.json; tail -3 gpurun_out/c_bench_c2.err
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/c_pytest.log 2>&1; tail -5 gpurun_out/c_pytest.log
