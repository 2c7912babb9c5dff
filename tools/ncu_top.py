"""Summarise an `ncu --page source --csv` export: per kernel section, the stall-reason breakdown and the top SASS
instructions by stall samples.  A report with several kernels has one table per kernel (each starts with an "Address"
header row; free-text rows in between name the kernel): round 1's version assumed a single table and crashed."""
import csv
import sys

path, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))


def section(title, hdr, data):
    col = {h: i for i, h in enumerate(hdr)}
    if "# Samples" not in col or not data:
        return
    S = col["# Samples"]

    def num(r, i):
        try:
            return int(float(r[i] or 0))
        except (ValueError, IndexError):
            return 0

    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(num(r, S) for r in data) or 1
    print(f"=== {title}: total samples {tot}")
    agg = {h: sum(num(r, col[h]) for r in data) for h in stall_cols}
    print("stall breakdown:", {k: f"{100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > 0.005 * tot})
    for r in sorted(data, key=lambda r: -num(r, S))[:topn]:
        n = num(r, S)
        top = sorted(((num(r, col[h]), h) for h in stall_cols), reverse=True)[:2]
        ex = r[col["Instructions Executed"]] if "Instructions Executed" in col and col["Instructions Executed"] < len(r) else "?"
        print(f"{100 * n / tot:5.2f}% {r[col['Address']][-5:]} exec={ex:>9} {r[col['Source']][:90]:90s} "
              f"{top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")


hdr, data, title, last_text = None, [], "kernel", ""
for r in rows:
    if r and r[0] == "Address":
        if hdr:
            section(title, hdr, data)
        hdr, data, title = r, [], last_text or "kernel"
    elif hdr and len(r) >= len(hdr):
        data.append(r)
    elif r:
        last_text = " ".join(x for x in r if x)[:120]
if hdr:
    section(title, hdr, data)
