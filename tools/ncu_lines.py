"""Top stalled SASS lines with their dominant stall reasons, from `ncu --page source --csv` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
iS, iSamp = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows if len(r) == len(hdr) and r[iSamp].isdigit()]
half = len(data) // 2 if len(data) > 4000 else len(data)   # the page lists the kernel twice when 2 launches match
data = data[:half]
tot = sum(int(r[iSamp]) for r in data)
agg = {hdr[i]: sum(int(r[i]) for r in data if r[i].isdigit()) for i in cols}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print(f"{k:28s} {100*v/tot:6.2f}%")
print("--- top lines")
order = sorted(range(len(data)), key=lambda n: -int(data[n][iSamp]))
for n in order[: int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    r = data[n]
    top = sorted(((int(r[i]) if r[i].isdigit() else 0, hdr[i][6:]) for i in cols), reverse=True)[:2]
    print(f"{100*int(r[iSamp])/tot:5.2f}% #{n:5d} {r[iS][:60]:60s} {top}")
