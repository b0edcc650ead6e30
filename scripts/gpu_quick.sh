#!/bin/bash
# quick measurement on the GPU box: per-pass times of c3 (and optionally c2) + in-bench parity; optional GPU tests
mkdir -p gpurun_out
TAG=${1:-quick}
python bench.py --skip-e2e --skip-cpu --steps 5 --warmup 2 > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
cat gpurun_out/${TAG}_c3.json; tail -3 gpurun_out/${TAG}_c3.err
if [ "$2" == "c2" ] || [ "$3" == "c2" ]; then
  python bench.py --config c2 --skip-e2e --skip-cpu --steps 5 --warmup 2 > gpurun_out/${TAG}_c2.json 2> gpurun_out/${TAG}_c2.err
  cat gpurun_out/${TAG}_c2.json; tail -3 gpurun_out/${TAG}_c2.err
fi
if [ "$2" == "tests" ] || [ "$3" == "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
fi
