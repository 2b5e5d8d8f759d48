"""sum an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (last N launches)"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
last = int(sys.argv[2]) if len(sys.argv) > 2 else 10 ** 9
i0 = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
h = rows[i0]
seq = []
for r in rows[i0 + 1:]:
    if len(r) < len(h):
        continue
    x = dict(zip(h, r))
    if x.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(x['Metric Value'].replace(',', ''))
    v = v / 1e3 if x['Metric Unit'] == 'ns' else v * 1e3 if x['Metric Unit'] == 'ms' else v
    seq.append((x['Kernel Name'][:70] + ' ' + x.get('Grid Size', ''), v))
d = collections.defaultdict(lambda: [0, 0.0])
for name, v in seq[-last:]:
    d[name][0] += 1
    d[name][1] += v
tot = sum(t for _, t in d.values())
print(f"total {tot:.1f} us over {sum(c for c, _ in d.values())} launches")
for k, (c, t) in sorted(d.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  {c:5d} x {t / c:8.1f}  {k}")
