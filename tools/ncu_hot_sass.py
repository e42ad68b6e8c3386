"""Top stall-sample instructions of an ncu source-page CSV (ncu -i X --page source --csv --print-source sass)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[isamp] or 0) for r in data)
print("total samples", tot)
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[c] or 0), hdr[c]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[isamp]):7d} {100*int(r[isamp])/tot:5.1f}%  ex={r[iex]:>9}  {r[isrc].strip()[:70]:70s} {st}")
