#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 400 -x > $O/pytest_e2e.log 2>&1; tail -n 5 $O/pytest_e2e.log
timeout 900 python bench.py --steps 10 --warmup 5 --no-gpu-reference --no-cpu-baseline > $O/bench_mseg3d.log 2>&1; tail -c 1800 $O/bench_mseg3d.log
