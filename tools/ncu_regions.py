"""Instruction / stall-sample shares of line ranges of one file from an `ncu --page source --csv --print-source sass,cuda` dump.
   python tools/ncu_regions.py dump.csv file.cu name:a-b name:a-b ..."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
fname = sys.argv[2]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]; c = {k: i for i, k in enumerate(h)}
cur = ""
byfile, sampf, lines = collections.Counter(), collections.Counter(), {}
for r in rows[hi + 1:]:
    if len(r) < len(h):
        if r and r[0] == "File Path": cur = r[1].split('/')[-1]
        continue
    if not r[0]: continue
    off = len(r) - len(h)          # unescaped quotes in asm lines add columns
    try:
        n = int(float(r[c["Instructions Executed"] + off] or 0)); s = int(float(r[c["# Samples"] + off] or 0))
    except ValueError:
        continue
    byfile[cur] += n; sampf[cur] += s
    if cur == fname: lines[int(r[0])] = (n, s)
tot, ts = sum(byfile.values()), sum(sampf.values())
print(f"total inst {tot:,} samples {ts:,}")
for f in byfile: print(f"  {f:28s} {byfile[f]/tot*100:5.1f}% inst {sampf[f]/ts*100:5.1f}% samp")
for spec in sys.argv[3:]:
    name, rng = spec.split(":"); a, b = map(int, rng.split("-"))
    n = sum(v[0] for k, v in lines.items() if a <= k <= b); s = sum(v[1] for k, v in lines.items() if a <= k <= b)
    print(f"{name:18s} {a:4d}-{b:<4d} inst {n/tot*100:5.1f}%  samp {s/ts*100:5.1f}%")
