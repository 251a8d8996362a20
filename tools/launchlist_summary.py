"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel name."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
t = collections.defaultdict(float)
n = collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    t[r[ki]] += v
    n[r[ki]] += 1
tot = sum(t.values())
print(f"total {tot / 1e3:.1f} ms in {sum(n.values())} launches")
for k, v in sorted(t.items(), key=lambda kv: -kv[1]):
    print(f"{v:11.1f} us {100 * v / tot:5.1f}%  n={n[k]:4d}  avg {v / n[k]:9.1f} us  {k[:150]}")
