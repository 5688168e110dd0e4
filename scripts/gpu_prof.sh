#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for P in 1 0; do
  LS3D_PRECISE=$P timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_gemm -o $O/prof_gemm_p$P -f \
      python scripts/prof_gemm.py > $O/ncu_gemm_p$P.log 2>&1
done
timeout 500 python bench.py --workload spconv_sweep --steps 5 --warmup 3 > $O/bench_sweep.log 2>&1
tail -n 5 $O/ncu_gemm_p1.log; tail -c 1500 $O/bench_sweep.log
