#!/bin/bash
# Round-2 session D: warp-aggregated mbarrier arrivals (all GEMM engines), split-weight camera convolutions.
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py -m gpu -q --timeout 120 -x > $O/pytest_kernels.log 2>&1; tail -n 8 $O/pytest_kernels.log
timeout 600 python scripts/diag_once.py > $O/diag_once.log 2>&1; tail -n 3 $O/diag_once.log | cut -c1-400
timeout 900 python scripts/diag_parity.py > $O/diag_parity.log 2>&1; tail -n 12 $O/diag_parity.log
timeout 900 python bench.py --steps 10 --warmup 5 --no-gpu-reference --no-parity --no-cpu-baseline > $O/bench_mseg3d.log 2>&1; tail -c 3000 $O/bench_mseg3d.log
