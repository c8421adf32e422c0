"""Per-source-line stall attribution from an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [launch_index] [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
idx = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--launch-skip", idx,
                      "--launch-count", "1", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, items = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 6 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].strip():      # a source-line row (SASS rows have an empty Line No)
        d = dict(zip(hdr, r))
        try:
            n = int(d["# Samples"])
        except ValueError:
            continue
        items.append((n, cur_file, r[0], r[1].strip(), d))
tot = sum(i[0] for i in items)
print(f"total samples {tot}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(i[4][s] or 0) for i in items) for s in stalls}
print("stall mix:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for n, f, ln, src, d in sorted(items, key=lambda i: -i[0])[:top]:
    st = sorted(((int(d[s] or 0), s[6:]) for s in stalls), reverse=True)[:3]
    print(f"{100 * n / tot:5.1f}% {f}:{ln:>4} inst={d['Instructions Executed']:>9} {src[:90]:<90} {st}")
