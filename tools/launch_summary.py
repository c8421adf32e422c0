import csv, collections, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
n = 0
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('<unnamed>::', '')
    val = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    val = val / 1e3 if unit == 'ns' else (val * 1e3 if unit == 'ms' else val)
    agg[name][0] += 1; agg[name][1] += val; n += 1
tot = sum(v[1] for v in agg.values())
print(f"# {path}: {n} launches, {tot/1e3:.2f} ms total (cold-cache, serialised: compare shares)")
print(f"{'us':>10} {'share':>6} {'n':>5} {'avg_us':>8}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{v[1]:10.1f} {100*v[1]/tot:5.1f}% {v[0]:5d} {v[1]/v[0]:8.1f}  {k[:80]}")
