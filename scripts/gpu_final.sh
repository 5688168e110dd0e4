#!/bin/bash
# Round-end style GPU session: parity tests, smoke, MSeg3D + SDSeg3D bench lines (no CPU baseline: that leg is timed by the
# driver's own run), conv scaling diagnostic.
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 700 python -m pytest tests -m gpu -q --timeout 200 > $O/pytest_gpu.log 2>&1; tail -n 4 $O/pytest_gpu.log
timeout 120 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
timeout 250 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_mseg3d.log 2>&1
timeout 250 python bench.py --workload sdseg3d_semantickitti --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_sdseg3d.log 2>&1
python - <<'PY'
import json
for f in ['bench_mseg3d','bench_sdseg3d']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.log').read().strip().splitlines()[-1])
        print(f, 'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), 'launches', d['gpu_launches'], {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k in ('achieved','frac','avg_launch_us','share_of_step','tflops')})
    except Exception as e:
        print(f, 'ERR', e, open(f'gpurun_out/{f}.log').read()[-1500:])
PY
