#!/bin/bash
# Round-2 session B: the gather-once engine under test (unit tests first, then everything), quick bench.
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q --timeout 120 > $O/pytest_gemm.log 2>&1; tail -n 25 $O/pytest_gemm.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 --deselect tests/test_gpu_gemm.py > $O/pytest_gpu.log 2>&1; tail -n 15 $O/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 5 --no-gpu-reference > $O/bench_mseg3d.log 2>&1; tail -c 4500 $O/bench_mseg3d.log
