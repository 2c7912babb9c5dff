"""SASS evidence: per kernel, counts of the Blackwell-native mnemonics (UTC*MMA = tcgen05.mma, LDTM/STTM =
tcgen05.ld/st, UTMALDG/UTMAREDG/UBLKCP = TMA, SYNCS = mbarrier) and of legacy tensor ops (HMMA must be 0)."""
import re
import subprocess
import sys
from collections import Counter

so = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur = None
stats = {}
for line in txt.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        stats[cur] = Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        stats[cur][m.group(2)] += 1
keys = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAREDG", "UBLKCP", "SYNCS", "MUFU", "FFMA2", "FADD2", "HMMA", "LDL", "STL"]
print("kernel".ljust(64), " ".join(k.rjust(8) for k in keys), " total")
for k, c in stats.items():
    print(k[:64].ljust(64), " ".join(str(c.get(x, 0)).rjust(8) for x in keys), str(sum(c.values())).rjust(6))
