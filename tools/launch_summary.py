"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('<unnamed>::', '')
    v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    if 'k_gemm' in name:
        name += ' narrow' if row['Grid Size'].strip('()').split(',')[0].strip() == '1' else ' big'
    agg[name][0] += 1; agg[name][1] += v; tot += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:60s} n={n:5d} total={t/1e3:10.3f} ms  {100*t/tot:5.1f}%")
print("total ms %.3f" % (tot / 1e3))
