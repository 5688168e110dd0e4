#!/bin/bash
# GPU tests + the model benches (no diagnostics): the cheap confirmation run.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 700 python -m pytest tests -m gpu -q --timeout 200 > $O/pytest_gpu.log 2>&1; tail -n 6 $O/pytest_gpu.log
timeout 250 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_mseg3d.log 2>&1
timeout 200 python bench.py --workload sdseg3d_semantickitti --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_sdseg3d.log 2>&1
python - <<'PY'
import json
for f in ['bench_mseg3d','bench_sdseg3d']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.log').read().strip().splitlines()[-1])
        print(f, 'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k in ('achieved','frac','avg_launch_us','share_of_step','tflops')}, d['roofline']['all_gemm'])
    except Exception as e:
        print(f, 'ERR', e, open(f'gpurun_out/{f}.log').read()[-1500:])
PY
