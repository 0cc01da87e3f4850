"""Per-kernel shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list (the summary committed beside the csv)."""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
hdr = [r for r in rows if r and r[0] == "ID"][0]
data = [r for r in rows if len(r) == len(hdr) and r[0].isdigit()]
iK, iM, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
agg = OrderedDict()
for r in data:
    if r[iM] != "gpu__time_duration.sum":
        continue
    name = r[iK].split("(")[0].replace("void ", "").replace("d2t::", "").replace("<unnamed>::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iV].replace(",", "")) / 1e3
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':62s} {'launches':>8s} {'total us':>10s} {'share':>7s} {'avg us':>8s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:62]:62s} {n:8d} {us:10.1f} {100 * us / tot:6.1f}% {us / n:8.1f}")
print(f"{'total':62s} {sum(a[0] for a in agg.values()):8d} {tot:10.1f}")
