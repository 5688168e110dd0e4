#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -x -k "conv" > $O/pytest_kernels.log 2>&1; tail -n 5 $O/pytest_kernels.log
timeout 600 python scripts/prof_camera.py > $O/prof_camera.log 2>&1; tail -n 45 $O/prof_camera.log
