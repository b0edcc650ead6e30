#!/bin/bash
mkdir -p gpurun_out
ldd multiview-reconstruction_b200/libmvdecon.so | grep -E "libz|not found"
( time timeout 400 python -m pytest tests/test_extras_gpu.py tests/test_abi.py -q ) > gpurun_out/sanity_pytest.log 2>&1; grep -E "passed|failed" gpurun_out/sanity_pytest.log | tail -2
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --steps 3 --warmup 3 --skip-cpu --e2e-iterations 3 2>/dev/null | tail -c 400
