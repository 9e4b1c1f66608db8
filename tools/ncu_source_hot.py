"""Summarise an `ncu --page source --csv --print-source sass` export: the instructions that collect the warp-stall samples, with the dominant
stall reason of each (authoring aid).  usage: ncu -i x.ncu-rep --page source --csv --print-source sass > x.csv; python tools/ncu_source_hot.py x.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = []
tot = 0
for k, r in enumerate(rows[hdr_i + 1:]):
    if len(r) < len(hdr):
        continue
    s = int(r[col["# Samples"]] or 0)
    tot += s
    data.append((k, r, s))
print("kernel:", rows[0][1] if rows[0] else "?", " total samples:", tot, " instructions:", len(data))
agg = {}
for k, r, s in data:
    for n in stalls:
        agg[n] = agg.get(n, 0) + int(r[col[n]] or 0)
print("stall totals:", ", ".join("%s %.1f%%" % (n[6:], 100.0 * v / max(1, tot)) for n, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
print("%5s %7s %6s  %-70s %s" % ("idx", "samples", "%", "instruction", "top stalls"))
for k, r, s in sorted(data, key=lambda x: -x[2])[:top]:
    st = sorted(((int(r[col[n]] or 0), n[6:]) for n in stalls), reverse=True)[:3]
    print("%5d %7d %5.1f%%  %-70s %s" % (k, s, 100.0 * s / max(1, tot), r[col["Source"]].strip()[:70], " ".join("%s:%d" % (n, v) for v, n in st if v)))
