#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -x -k "conv or pad3 or upsample" > $O/pytest_kernels.log 2>&1; tail -n 15 $O/pytest_kernels.log
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 400 -x > $O/pytest_e2e.log 2>&1; tail -n 15 $O/pytest_e2e.log
timeout 600 python scripts/prof_camera.py > $O/prof_camera.log 2>&1; tail -n 42 $O/prof_camera.log
CAM_MODE=fp16 timeout 600 python scripts/prof_camera.py > $O/prof_camera_fp16.log 2>&1; head -n 14 $O/prof_camera_fp16.log

timeout 900 python bench.py --steps 10 --warmup 5 --no-gpu-reference --no-cpu-baseline > $O/bench_mseg3d.log 2>&1; tail -c 1300 $O/bench_mseg3d.log
