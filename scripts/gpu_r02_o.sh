#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -x -k "sffm_decoder" > $O/pytest_dec.log 2>&1; tail -n 15 $O/pytest_dec.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -x -k "head_vs_reference or class_embed" -rs > $O/pytest_head.log 2>&1; tail -n 8 $O/pytest_head.log
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 400 -x > $O/pytest_e2e.log 2>&1; tail -n 5 $O/pytest_e2e.log
timeout 900 python bench.py --steps 10 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-secondary --no-parity > $O/bench_fused.log 2>&1; tail -c 1500 $O/bench_fused.log
LS3D_FUSED_DECODER=0 timeout 900 python bench.py --steps 10 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-secondary --no-parity > $O/bench_unfused.log 2>&1; tail -c 600 $O/bench_unfused.log
