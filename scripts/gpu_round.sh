#!/bin/bash
# One GPU session: parity tests, diagnostics, bench lines, ncu launch list + one full capture of the dominant kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "from lidarseg3d_b200 import capi; capi.lib(); print('lib ok')" > $O/lib.log 2>&1
timeout 300 python scripts/diag_gemm.py > $O/diag_gemm.log 2>&1
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x > $O/pytest_gpu.log 2>&1
timeout 400 python scripts/diag_e2e.py > $O/diag_e2e.log 2>&1
timeout 120 python __graft_entry__.py --smoke > $O/smoke.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_mseg3d.log 2>&1
timeout 300 python bench.py --workload sdseg3d_semantickitti --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_sdseg3d.log 2>&1
if [ "$1" == "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_mseg3d.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_gemm -s 120 -c 3 -o $O/prof_gather_gemm \
      python bench.py --workload sdseg3d_semantickitti --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
fi
tail -n 3 $O/lib.log $O/pytest_gpu.log $O/smoke.log
tail -c 600 $O/bench_mseg3d.log
