#!/bin/bash
# Round-2 session A: parity tests, smoke, bench lines (with parity / CPU / GPU-reference blocks), launch list and a
# `--set full` capture of the TIMED gather-GEMM engine inside a real forward.
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -x > $O/pytest_gpu.log 2>&1; tail -n 15 $O/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_mseg3d.log 2>&1; tail -c 6000 $O/bench_mseg3d.log
timeout 600 python bench.py --workload sdseg3d_semantickitti --steps 20 --warmup 5 --no-parity --no-cpu-baseline > $O/bench_sdseg3d.log 2>&1; tail -c 1500 $O/bench_sdseg3d.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gather_gemm -c 16 \
    -o $O/prof_unet -f python scripts/prof_unet_step.py > $O/ncu_unet.log 2>&1; tail -n 3 $O/ncu_unet.log
ncu -i $O/prof_unet.ncu-rep --page raw --csv > $O/prof_unet_raw.csv 2>/dev/null
SZ=$(stat -c %s $O/prof_unet.ncu-rep 2>/dev/null || echo 0); if [ "$SZ" -gt 40000000 ]; then rm -f $O/prof_unet.ncu-rep; fi
LS3D_PROFILE_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
    --log-file $O/launches_mseg3d.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-gpu-reference --no-secondary --eager-images > $O/ncu_bench.log 2>&1
tail -n 2 $O/ncu_bench.log | cut -c1-300
