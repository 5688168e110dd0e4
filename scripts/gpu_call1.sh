#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.log 2>&1
timeout 300 python scripts/diag_skip.py > $O/diag_skip.log 2>&1; cat $O/diag_skip.log | tail -n 12
timeout 700 python -m pytest tests -m gpu -q --timeout 200 > $O/pytest_gpu.log 2>&1; tail -n 6 $O/pytest_gpu.log
timeout 250 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_mseg3d.log 2>&1; tail -c 1500 $O/bench_mseg3d.log
