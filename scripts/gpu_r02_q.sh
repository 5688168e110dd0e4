#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -x -k "sffm_decoder or head_vs_reference" > $O/pytest_dec.log 2>&1; tail -n 4 $O/pytest_dec.log
timeout 900 python bench.py --steps 10 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-secondary --no-parity > $O/bench_q.log 2>&1; tail -c 2500 $O/bench_q.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sffm_decoder -c 4 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "sffm_decoder" 2>&1 | grep -E "sffm_decoder_kernel|gpu__time" | head -12
