"""Top stall-sample instructions of an ncu source-page CSV (ncu -i X --page source --csv --print-source sass
[--kernel-name regex:...]); the first kernel instance of the file.

    python tools/ncu_hot_sass.py source.csv [top_n]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
print(rows[0][1] if rows and len(rows[0]) > 1 else "")
hdr = rows[1]
end = len(rows)
for i in range(2, len(rows)):
    if rows[i] and rows[i][0] == "Kernel Name":
        end = i
        break
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:end] if len(r) == len(hdr)]
tot = sum(int(r[isamp] or 0) for r in data)
print("instructions", len(data), "total samples", tot, "warp instructions executed", sum(int(r[iex] or 0) for r in data))
agg = {}
for r in data:
    for c in stall_cols:
        agg[hdr[c]] = agg.get(hdr[c], 0) + int(r[c] or 0)
print("stall reasons:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[c] or 0), hdr[c]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[isamp]):7d} {100*int(r[isamp])/max(tot,1):5.1f}%  ex={r[iex]:>9}  {r[isrc].strip()[:70]:70s} {st}")
