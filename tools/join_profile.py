"""Join tools/step_profile.py's op log with the ncu launch list of the same run.
usage: python tools/join_profile.py step_ops.json step_launches.csv [--top N] [--peak GBps]
Prints (1) per-kernel-name totals and (2) the per-call table sorted by time with the GB/s each call achieves on the bytes of
the tensors it touches (an upper bound on its algorithmic traffic) and the fraction of the measured HBM peak."""
import collections
import csv
import json
import re
import sys

ops = json.load(open(sys.argv[1]))
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
peak = float(sys.argv[sys.argv.index("--peak") + 1]) if "--peak" in sys.argv else 6527.1
with open(sys.argv[2]) as f:
    lines = [l for l in f if not l.startswith("==")]
kern = []
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    kern.append((name, us))
# logs written before step_profile.py wrapped OptimState.advance() lack that op: drop its kernel so the positional join holds
have_adv = any(o["op"] == "optim_advance" for o in ops)
ours = [(n, t) for n, t in kern if "at::" not in n and "elementwise" not in n and "Memset" not in n
        and (have_adv or "optim_advance" not in n)]
ops = [o for o in ops if o["op"] not in ("bn_fin", "bn_bwd_fin", "se_bn")]
need = sum(o["launches"] for o in ops)
print(f"# {len(kern)} launches in the csv, {len(ours)} from libteethrt, op log expects {need}")
if len(ours) != need:
    print("# WARNING: counts differ; the join below is positional and may be shifted")
tot = sum(t for _, t in kern)
print(f"# total kernel time {tot / 1e3:.3f} ms (serialised under ncu); torch-side kernels {sum(t for n, t in kern if (n, t) not in ours) / 1e3:.3f} ms")
rows, i = [], 0
for o in ops:
    ks = ours[i:i + o["launches"]]
    i += o["launches"]
    t = sum(x[1] for x in ks)
    rows.append(dict(op=o["op"], us=t, bytes=o["bytes"], kernels=[k[0] for k in ks], shapes=o["shapes"]))
agg = collections.defaultdict(lambda: [0, 0.0, 0])
for r in rows:
    a = agg[r["op"]]
    a[0] += 1
    a[1] += r["us"]
    a[2] += r["bytes"]
print(f"\n{'op':<18}{'calls':>6}{'us':>10}{'share':>7}{'GB':>8}{'GB/s':>8}{'hbm%':>6}{'us@peak':>9}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = a[2] / max(a[1], 1e-9) / 1e3
    print(f"{k:<18}{a[0]:>6}{a[1]:>10.1f}{100 * a[1] / tot:>6.1f}%{a[2] / 1e9:>8.2f}{gbs:>8.0f}{100 * gbs / peak:>6.1f}{a[2] / peak / 1e3:>9.1f}")
print(f"{'sum':<18}{len(rows):>6}{sum(r['us'] for r in rows):>10.1f}{'':>7}{sum(r['bytes'] for r in rows) / 1e9:>8.2f}"
      f"{'':>14}{sum(r['bytes'] for r in rows) / peak / 1e3:>9.1f}")
print(f"\n{'op':<16}{'us':>9}{'MB':>9}{'GB/s':>7}{'hbm%':>6}  shapes")
for r in sorted(rows, key=lambda r: -r["us"])[:top]:
    gbs = r["bytes"] / max(r["us"], 1e-9) / 1e3
    shp = " ".join("x".join(map(str, s)) for s in r["shapes"][:5])
    print(f"{r['op']:<16}{r['us']:>9.1f}{r['bytes'] / 1e6:>9.1f}{gbs:>7.0f}{100 * gbs / peak:>6.1f}  {shp}")
