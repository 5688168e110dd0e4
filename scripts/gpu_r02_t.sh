#!/bin/bash
# ncu --set full of the timed sparse engine inside a real forward -> traffic json; SDSeg3D / Waymo lines; full sparse-conv sweep
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gather_gemm -c 16 \
    -o $O/prof_unet -f python scripts/prof_unet_step.py > $O/ncu_unet.log 2>&1; tail -n 3 $O/ncu_unet.log
ncu -i $O/prof_unet.ncu-rep --page raw --csv > $O/prof_unet_raw.csv 2>/dev/null
python scripts/ncu_traffic.py $O/prof_unet_raw.csv $O/unet_launches.json $O/r02_gather_gemm_traffic.json "round-2 gather-once engine, one eager MSeg3D forward of the bench batch"
SZ=$(stat -c %s $O/prof_unet.ncu-rep 2>/dev/null || echo 0); if [ "$SZ" -gt 30000000 ]; then rm -f $O/prof_unet.ncu-rep; fi
timeout 600 python bench.py --workload sdseg3d_semantickitti --steps 20 --warmup 5 > $O/bench_sdseg3d.log 2>&1; tail -c 1200 $O/bench_sdseg3d.log
timeout 600 python bench.py --workload mseg3d_waymo --steps 10 --warmup 5 --no-secondary --no-gpu-reference > $O/bench_waymo.log 2>&1; tail -c 1200 $O/bench_waymo.log
timeout 1200 python bench.py --workload spconv_sweep --sweep-full --steps 3 --warmup 1 > $O/bench_sweep.log 2>&1; tail -c 3000 $O/bench_sweep.log
