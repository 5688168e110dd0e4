#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/trace_once.py > $O/trace_once.log 2>&1; tail -n 3 $O/trace_once.log
