"""Print an `ncu --csv --metrics ...` launch list as one line per launch."""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
hdr = [r for r in rows if r and r[0] == "ID"][0]
data = [r for r in rows if len(r) == len(hdr) and r[0].isdigit()]
iK, iM, iV, iG = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
d = OrderedDict()
for r in data:
    d.setdefault(r[0], {"k": r[iK].split("(")[0].replace("void ", "").replace("d2t::", ""), "g": r[iG]})[r[iM]] = r[iV]
tot = 0.0
for k, v in d.items():
    ns = float(v.get("gpu__time_duration.sum", "0").replace(",", ""))
    tot += ns
    rest = {m.split(".")[0].replace("smsp__", ""): x for m, x in v.items() if m not in ("k", "g", "gpu__time_duration.sum")}
    print(f"{int(k):4d} {v['k'][:40]:40s} {v['g']:>16s} {ns/1e3:9.1f} us  {rest}")
print(f"total {tot/1e3:.1f} us")
