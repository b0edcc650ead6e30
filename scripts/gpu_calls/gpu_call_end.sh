#!/bin/bash
mkdir -p gpurun_out
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 40 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e 2>/dev/null | tail -c 300
