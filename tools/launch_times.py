"""Average per-kernel duration from an ncu launch-list CSV (gpu__time_duration.sum)."""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
acc = OrderedDict()
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    name = r[ik].split("(")[0].replace("void ", "")[:40]
    acc.setdefault(name, []).append(float(r[iv].replace(",", "")))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 1
tot = 0.0
for k, v in acc.items():
    v = v[skip:] if len(v) > skip else v
    m = sum(v) / len(v)
    if "synth" not in k:
        tot += m
    print(f"  {k:42s} n={len(v):3d} avg={m/1000:9.2f} us")
print(f"  sum (excluding synth) = {tot/1000:.2f} us")
