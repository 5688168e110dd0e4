#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 100 -k "class_embed or head_vs_reference or sffm" > $O/pytest_x.log 2>&1; tail -n 3 $O/pytest_x.log
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 400 -x > $O/pytest_e2e.log 2>&1; tail -n 3 $O/pytest_e2e.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-parity > $O/bench_x.log 2>&1; tail -c 300 $O/bench_x.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:class_tokens -c 2 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "head_vs_reference or class_embed" 2>&1 | grep -E "class_tokens_kernel|gpu__time" | head -6
