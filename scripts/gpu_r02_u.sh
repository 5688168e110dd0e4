#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_inputs.py -m gpu -q --timeout 100 -k "upsample or resize or conv_plan" > $O/pytest_u.log 2>&1; tail -n 5 $O/pytest_u.log
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 400 -x > $O/pytest_e2e.log 2>&1; tail -n 4 $O/pytest_e2e.log
timeout 900 python bench.py --steps 10 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-secondary --no-parity > $O/bench_u.log 2>&1; tail -c 400 $O/bench_u.log
LS3D_PROFILE_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
    --log-file $O/launches_mseg3d.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-gpu-reference --no-secondary --eager-images > $O/ncu_bench.log 2>&1
python scripts/summarize_launches.py $O/launches_mseg3d.csv > $O/launches_summary.txt; head -24 $O/launches_summary.txt | cut -c1-150
