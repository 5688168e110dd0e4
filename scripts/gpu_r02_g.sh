#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -x -k "conv or pad3 or upsample" > $O/pytest_kernels.log 2>&1; tail -n 25 $O/pytest_kernels.log
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 400 -x -k "dual" > $O/pytest_e2e.log 2>&1; tail -n 25 $O/pytest_e2e.log
timeout 900 python scripts/diag_parity.py > $O/diag_parity.log 2>&1; tail -n 8 $O/diag_parity.log | cut -c1-700
