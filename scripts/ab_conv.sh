#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
echo "== A (in-tree)"; timeout 100 python -m pytest tests/test_gpu_kernels.py -q -k "conv3x3" --timeout 40 2>&1 | tail -2; timeout 100 python scripts/diag_conv.py 2>&1 | grep timing | cut -c1-75
cp scripts/variants/conv3x3_v8.cu.txt lidarseg3d_b200/csrc/conv3x3_f16.cu
python -c "from lidarseg3d_b200 import build; build.build()" 2>&1 | grep -v deprecated | tail -3
echo "== B (v8)"; timeout 100 python -m pytest tests/test_gpu_kernels.py -q -k "conv3x3" --timeout 40 2>&1 | tail -2; timeout 100 python scripts/diag_conv.py 2>&1 | grep timing | cut -c1-75
