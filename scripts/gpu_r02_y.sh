#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
LS3D_NO_CLOCKS=1 timeout 900 python bench.py --steps 30 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-parity --no-secondary > $O/bench_y0.log 2>&1; tail -c 200 $O/bench_y0.log
timeout 900 python bench.py --steps 30 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-parity --no-secondary > $O/bench_y1.log 2>&1; tail -c 200 $O/bench_y1.log
