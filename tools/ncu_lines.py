"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump by CUDA source line:
instructions executed, stall samples and the dominant stall reasons.   python tools/ncu_lines.py dump.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
c = {k: i for i, k in enumerate(h)}
src_col = 1
inst, samp = collections.Counter(), collections.Counter()
stalls = collections.defaultdict(collections.Counter)
text = {}
cur_file = ""
names = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
for r in rows[hi + 1:]:
    if len(r) < len(h):
        if r and r[0] == "File Path": cur_file = r[1].split("/")[-1]
        continue
    if not r[0]:           # SASS rows follow the aggregate row of their CUDA line
        continue
    key = (cur_file, r[0])
    try:
        n = int(float(r[c["Instructions Executed"]] or 0)); s = int(float(r[c["# Samples"]] or 0))
    except ValueError:
        continue
    inst[key] += n; samp[key] += s
    text.setdefault(key, r[src_col][:90])
    for k in names:
        v = r[c[k]]
        if v and v != "0": stalls[key][k[6:]] += int(float(v))
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {ti:,}  samples {ts:,}")
for key, s in samp.most_common(top):
    st = ", ".join(f"{k} {v}" for k, v in stalls[key].most_common(3))
    print(f"{key[0]}:{key[1]:>4} inst {inst[key]/ti*100:5.1f}% samp {s/ts*100:5.1f}%  [{st}]  {text[key]}")
