"""Summarise an `ncu --page source --csv` export: top SASS instructions by stall samples, grouped totals."""
import csv
import sys

path, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
data = rows[hdr_i + 1:]
S = col["# Samples"]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S] or 0) for r in data)
print("total samples", tot)
agg = {h: sum(int(r[col[h]] or 0) for r in data) for h in stall_cols}
print("stall breakdown:", {k: f"{100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > 0.005 * tot})
data_s = sorted(data, key=lambda r: -int(r[S] or 0))
for r in data_s[:topn]:
    n = int(r[S] or 0)
    top = sorted(((int(r[col[h]] or 0), h) for h in stall_cols), reverse=True)[:2]
    print(f"{100*n/tot:5.2f}% {r[col['Address']][-5:]} exec={r[col['Instructions Executed']]:>9} {r[col['Source']][:90]:90s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
