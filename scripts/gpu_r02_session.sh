#!/bin/bash
# Round-2 session M (re-entry): full GPU test suite, smoke, the default bench line exactly as the driver runs it, launch list.
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 > $O/pytest_gpu.log 2>&1; tail -n 15 $O/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
( time timeout 1200 python bench.py ) > $O/bench_mseg3d.log 2>&1; tail -c 9000 $O/bench_mseg3d.log
LS3D_PROFILE_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
    --log-file $O/launches_mseg3d.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-gpu-reference --no-secondary --eager-images > $O/ncu_bench.log 2>&1
tail -n 2 $O/ncu_bench.log | cut -c1-300
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/launches_mseg3d.csv') if l.startswith('"')))
hdr = rows[0]; ni = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        agg[r[ni][:100]][0] += 1; agg[r[ni][:100]][1] += float(r[vi].replace(',', ''))
    except Exception:
        pass
tot = sum(v[1] for v in agg.values())
out = [f'launches {sum(v[0] for v in agg.values())} total {tot/1e3:.1f} us (serialised, cold) -> {tot/1e6:.2f} ms per step']
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    out.append(f'{v[1]/1e3:10.1f} us {v[1]/tot*100:6.2f}%  n={v[0]:5d}  {k}')
open('gpurun_out/launches_summary.txt', 'w').write('\n'.join(out) + '\n')
print('\n'.join(out))
PY
