"""Instructions executed and stall samples per CUDA source line, from `ncu -i rep --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_srclines.py file.csv <kernel substring> [top N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cur_fn, cur_file, hdr, agg = None, None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        cur_fn = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and cur_fn and want in cur_fn and r[0].isdigit():
        iI, iS = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        key = (cur_fn[:50], cur_file, int(r[0]), r[1][:90])
        a = agg.setdefault(key, [0, 0])
        a[0] += int(r[iI]) if r[iI].isdigit() else 0
        a[1] += int(r[iS]) if r[iS].isdigit() else 0
fns = sorted({k[0] for k in agg})
for fn in fns:
    items = [(k, v) for k, v in agg.items() if k[0] == fn]
    ti, ts = sum(v[0] for _, v in items), sum(v[1] for _, v in items)
    print(f"== {fn}: {ti} warp instructions, {ts} samples")
    for k, v in sorted(items, key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * v[0] / max(ti, 1):5.1f}% inst {100 * v[1] / max(ts, 1):5.1f}% smpl  {k[1]}:{k[2]:4d}  {k[3]}")
