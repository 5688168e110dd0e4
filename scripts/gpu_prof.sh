#!/bin/bash
# ncu evidence for the dominant kernel: (1) launch list of one bench step, (2) full-set capture of three representative
# gather-GEMM launches (isolated script), exported to text/CSV; the .ncu-rep is kept only if small.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/*.ncu-rep
LS3D_PROFILE_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
    --log-file $O/launches_mseg3d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager-images > $O/ncu_bench.log 2>&1
LS3D_PRECISE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gather_gemm \
    -o $O/prof_gemm -f python scripts/prof_gemm.py > $O/ncu_gemm.log 2>&1
ncu -i $O/prof_gemm.ncu-rep --page raw --csv > $O/prof_gemm_raw.csv 2>/dev/null
ncu -i $O/prof_gemm.ncu-rep --page details > $O/prof_gemm_details.txt 2>/dev/null
ncu -i $O/prof_gemm.ncu-rep --page source --csv > $O/prof_gemm_source.csv 2>/dev/null
SZ=$(stat -c %s $O/prof_gemm.ncu-rep); if [ "$SZ" -gt 30000000 ]; then rm -f $O/prof_gemm.ncu-rep; fi
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/launches_mseg3d.csv') if l.startswith('"')))
hdr = rows[0]; ni = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        agg[r[ni][:70]][0] += 1; agg[r[ni][:70]][1] += float(r[vi].replace(',', ''))
    except Exception:
        pass
tot = sum(v[1] for v in agg.values())
print('launches', sum(v[0] for v in agg.values()), 'total_ns', tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f'{v[1]/tot*100:6.2f}%  {v[0]:5d}  {v[1]/1e3:10.1f} us  {k}')
PY


timeout 500 python bench.py --steps 10 --warmup 3 > $O/bench_default_with_cpu.log 2>&1; tail -c 1500 $O/bench_default_with_cpu.log
timeout 400 python bench.py --workload spconv_sweep --steps 5 --warmup 3 > $O/bench_sweep.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_sweep.log').read().strip().splitlines()[-1])
print('sweep value', d['value'], d['roofline'])
for r in d['sweep']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})
PY
