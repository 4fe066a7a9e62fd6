"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list per kernel.
usage: python profiles/launch_summary.py <csv>"""
import collections, csv, sys
lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
agg = collections.OrderedDict()
for r in rows:
    if r.get('Metric Name', 'gpu__time_duration.sum') != 'gpu__time_duration.sum': continue
    k = r['Kernel Name'].split('(')[0]
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += float(r['Metric Value'].replace(',', '')) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"{sum(v[0] for v in agg.values())} launches, {tot:.2f} ms in total (per-launch times under ncu are cold-cache and serialised)")
for k, v in agg.items():
    print(f"{k[:56]:56s} n={v[0]:4d}  total={v[1]:9.3f} ms  avg={v[1] / v[0]:8.4f} ms  share={v[1] / tot * 100:5.1f}%")
