"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py file.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[ix['Metric Name']] != 'gpu__time_duration.sum': continue
    name = r[ix['Kernel Name']].split('(')[0]
    v = float(r[ix['Metric Value']].replace(',', ''))
    mult = {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(r[ix['Metric Unit']], 1)
    agg[name][0] += 1; agg[name][1] += v * mult
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{k[:70]:70s} n={v[0]:5d} total={v[1]/1e3:10.3f} ms share={100*v[1]/tot:6.2f}% avg={v[1]/v[0]:9.1f} us")
print('total ms', tot / 1e3)
