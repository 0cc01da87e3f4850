"""Instruction mix / stall samples / shared wavefronts per opcode from `ncu --page source --csv` output."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
iSamp = hdr.index("Warp Stall Sampling (All Samples)")
iW, iWi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
tot, samp, wf, wfi = (collections.Counter() for _ in range(4))
for r in rows:
    if len(r) != len(hdr) or not r[iE].isdigit():
        continue
    toks = r[iS].split()
    if not toks:
        continue
    t = toks[1] if toks[0].startswith("@") else toks[0]
    op = t.split(".")[0] + (".128" if ".128" in t else "")
    n = int(r[iE])
    tot[op] += n
    samp[op] += int(r[iSamp])
    wf[op] += int(r[iW])
    wfi[op] += int(r[iWi])
T, S = sum(tot.values()), max(1, sum(samp.values()))
print("total warp-instr", T, "samples", S)
for op, n in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print(f"{op:12s} {n:12d} {100*n/T:6.2f}%  samples {100*samp[op]/S:6.2f}%  smem wavefronts {wf[op]:12d} ideal {wfi[op]:12d}")
