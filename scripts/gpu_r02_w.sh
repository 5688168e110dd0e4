#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q --timeout 100 > $O/pytest_w.log 2>&1; tail -n 5 $O/pytest_w.log
timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_kernels.py -m gpu -q --timeout 400 -k "e2e or unet or spmiddle or rulebook or gather" > $O/pytest_e2e.log 2>&1; tail -n 4 $O/pytest_e2e.log
timeout 900 python bench.py --steps 10 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-secondary --no-parity > $O/bench_w.log 2>&1; tail -c 300 $O/bench_w.log
timeout 300 python scripts/trace_once.py > $O/trace_once.log 2>&1; tail -n 1 $O/trace_once.log | cut -c1-100
