#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 600 python scripts/prof_camera.py > $O/prof_camera.log 2>&1; tail -n 42 $O/prof_camera.log
CAM_MODE=fp16 timeout 600 python scripts/prof_camera.py > $O/prof_camera_fp16.log 2>&1; head -n 30 $O/prof_camera_fp16.log
