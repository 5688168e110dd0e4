#!/bin/bash
# GPU session: all parity tests (incl. the fp16 camera-branch kernels), then MSeg3D bench lines fp32 vs fp16 camera branch.
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 700 python -m pytest tests -m gpu -q --timeout 200 > $O/pytest_gpu.log 2>&1; tail -n 25 $O/pytest_gpu.log

timeout 250 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --image-dtype fp16 > $O/bench_mseg3d_fp16.log 2>&1
python - <<'PY'
import json
for f in ['bench_mseg3d_fp16']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.log').read().strip().splitlines()[-1])
        print(f, 'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e, open(f'gpurun_out/{f}.log').read()[-1500:])
PY
