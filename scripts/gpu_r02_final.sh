#!/bin/bash
# Round-2 final session: full GPU test suite, smoke, the default bench line exactly as the driver runs it, reference arm.
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 > $O/pytest_gpu.log 2>&1; tail -n 6 $O/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
( time timeout 1200 python bench.py ) > $O/bench_mseg3d.log 2>&1; tail -c 3000 $O/bench_mseg3d.log
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $O/bench_reference.log 2>&1; tail -c 1200 $O/bench_reference.log
LS3D_PROFILE_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
    --log-file $O/launches_mseg3d.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-gpu-reference --no-secondary --eager-images > $O/ncu_bench.log 2>&1
python scripts/summarize_launches.py $O/launches_mseg3d.csv > $O/launches_summary.txt; head -16 $O/launches_summary.txt | cut -c1-150
