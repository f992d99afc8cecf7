"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel name count, total, mean; and the
per-level sequence of one build."""
import csv, sys, re, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr = r; rows = rows[i + 1:]; break
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
seq = []
for r in rows:
    if len(r) <= mv: continue
    name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("orb::", "")
    try: us = float(r[mv].replace(",", "")) / 1e3
    except ValueError: continue
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
    seq.append((name, us))
tot = sum(a[1] for a in agg.values())
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:55s} n={n:4d} total={us:9.1f} us  mean={us/n:7.1f} us  share={us/tot*100:5.1f}%")
print("total", round(tot, 1), "us")
if len(sys.argv) > 2:
    for name, us in seq[: int(sys.argv[2])]:
        print(f"   {name:55s} {us:8.1f}")
