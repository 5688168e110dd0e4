#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/prof_camera.py > $O/prof_camera.log 2>&1; head -n 3 $O/prof_camera.log
timeout 400 python bench.py --workload sdseg3d_semantickitti --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-parity > $O/bench_sdseg3d2.log 2>&1; tail -c 300 $O/bench_sdseg3d2.log
