#!/bin/bash
# Round-2 session C: dual-mode kernels + once-engine fallback under test; parity diagnostic; once-engine time breakdown.
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py -m gpu -q --timeout 120 -x > $O/pytest_kernels.log 2>&1; tail -n 8 $O/pytest_kernels.log
timeout 900 python scripts/diag_parity.py > $O/diag_parity.log 2>&1; tail -n 12 $O/diag_parity.log
timeout 600 python scripts/diag_once.py > $O/diag_once.log 2>&1; tail -n 3 $O/diag_once.log | cut -c1-600
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 400 > $O/pytest_e2e.log 2>&1; tail -n 8 $O/pytest_e2e.log
