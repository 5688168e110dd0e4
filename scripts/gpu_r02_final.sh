#!/bin/bash
# Round-2 final session: full GPU test suite, smoke, the default bench line exactly as the driver runs it, reference arm.
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 > $O/pytest_gpu.log 2>&1; tail -n 6 $O/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
( time timeout 1200 python bench.py ) > $O/bench_mseg3d.log 2>&1; tail -c 3000 $O/bench_mseg3d.log
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $O/bench_reference.log 2>&1; tail -c 1200 $O/bench_reference.log
