"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (argv[1])."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ni, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        agg[r[ni][:100]][0] += 1
        agg[r[ni][:100]][1] += float(r[vi].replace(',', ''))
    except Exception:
        pass
tot = sum(v[1] for v in agg.values())
print(f'launches {sum(v[0] for v in agg.values())} total {tot / 1e3:.1f} us (serialised, cold) -> {tot / 1e6:.2f} ms per step')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f'{v[1] / 1e3:10.1f} us {v[1] / tot * 100:6.2f}%  n={v[0]:5d}  {k}')
