"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (+ grid)."""
import collections
import csv
import re
import sys

rows = collections.OrderedDict()
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, ig, iv = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
iu = hdr.index("Metric Unit")
for r in rd:
    if len(r) <= iv:
        continue
    name = re.sub(r"\(.*", "", r[ik])[:70]
    t = float(r[iv].replace(",", ""))
    if r[iu] in ("nsecond", "ns"):
        t /= 1000.0
    elif r[iu] in ("msecond", "ms"):
        t *= 1000.0
    key = (name, r[ig])
    n, s = rows.get(key, (0, 0.0))
    rows[key] = (n + 1, s + t)
tot = sum(s for _, s in rows.values())
print("total %.1f us over %d launches" % (tot, sum(n for n, _ in rows.values())))
print("| kernel | grid | launches | total us | avg us | share |\n|---|---|---|---|---|---|")
for (name, grid), (n, s) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("| `%s` | %s | %d | %.0f | %.1f | %.1f %% |" % (name, grid, n, s, s / n, 100 * s / tot))
