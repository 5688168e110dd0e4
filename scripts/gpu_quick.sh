#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/*.ncu-rep
timeout 150 python scripts/diag_gemm.py > $O/diag_gemm.log 2>&1
RC=$?
grep -E "^===|^[a-z_0-9]+ [0-9.e-]+$|EXC|timing" $O/diag_gemm.log | head -60
if [ $RC -ne 0 ]; then echo "diag_gemm failed rc=$RC - stopping"; tail -n 20 $O/diag_gemm.log; exit 0; fi
timeout 600 python -m pytest tests -m gpu -q --timeout 150 -x > $O/pytest_gpu.log 2>&1; tail -n 4 $O/pytest_gpu.log
timeout 250 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_mseg3d.log 2>&1
timeout 200 python bench.py --workload sdseg3d_semantickitti --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_sdseg3d.log 2>&1
timeout 300 python scripts/diag_e2e.py > $O/diag_e2e.log 2>&1; grep -E '^\{|cpu launch' $O/diag_e2e.log; sed -n '/Self CPU %/,$p' $O/diag_e2e.log | tail -n 40 | cut -c1-60,150-250
python - <<'PY'
import json
for f in ['bench_mseg3d','bench_sdseg3d']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.log').read().strip().splitlines()[-1])
        print(f, 'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k in ('achieved','frac','avg_launch_us','share_of_step','tflops')}, d['roofline']['all_gemm'])
    except Exception as e:
        print(f, 'ERR', e, open(f'gpurun_out/{f}.log').read()[-800:])
PY
