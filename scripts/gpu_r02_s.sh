#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm.py -m gpu -q --timeout 100 -k "conv_kb or pack_bf16x3 or spmiddle or conv_plan" > $O/pytest_kb.log 2>&1; tail -n 12 $O/pytest_kb.log
timeout 600 python scripts/prof_camera.py > $O/prof_camera.log 2>&1; head -n 24 $O/prof_camera.log
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 400 -x > $O/pytest_e2e.log 2>&1; tail -n 4 $O/pytest_e2e.log
timeout 900 python bench.py --steps 10 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-secondary > $O/bench_s.log 2>&1; tail -c 1500 $O/bench_s.log
