#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
for nh in 1 2 4; do for lazy in 0 1; do
echo "== nh=$nh lazy=$lazy"; LS3D_C3_NH=$nh LS3D_C3_LAZY=$lazy timeout 60 python scripts/diag_conv.py 2>&1 | grep timing | cut -c1-70
done; done
