#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py --steps 10 --warmup 5 --no-gpu-reference --no-cpu-baseline > $O/bench_mseg3d.log 2>&1; tail -c 2500 $O/bench_mseg3d.log
LS3D_PROFILE_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
    --log-file $O/launches_mseg3d_dual.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-gpu-reference --no-secondary --eager-images > $O/ncu_bench.log 2>&1
tail -n 2 $O/ncu_bench.log | cut -c1-300
