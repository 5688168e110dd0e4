#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
for G in 4 2 1; do
  echo "== group $G"
  LS3D_ONCE_GROUP=$G timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q --timeout 120 -x 2>&1 | tail -n 3
  LS3D_ONCE_GROUP=$G DIAG_FAST=1 timeout 300 python scripts/diag_once.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        r = json.loads(l)
        print(r['name'], r['cin'], r['cout'], 'once %.1f pair %.1f' % (r['once_us'], r['pair_us']))
"
done
