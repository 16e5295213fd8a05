"""Aggregate an `ncu --page source --csv` export by code region: regions are delimited by marker opcodes.
usage: python profiles/ncu_regions.py file.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
base = int(data[0][ix["Address"]], 16)
tot_s = sum(int(r[ix["# Samples"]]) for r in data)
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)
# regions: split whenever executed-count changes by >30% AND opcode class changes: simpler -- fixed windows of 64 instrs
W = int(sys.argv[2]) if len(sys.argv) > 2 else 128
print("total samples", tot_s, "instructions executed", tot_i)
for s in range(0, len(data), W):
    chunk = data[s:s + W]
    smp = sum(int(r[ix["# Samples"]]) for r in chunk)
    ins = sum(int(r[ix["Instructions Executed"]]) for r in chunk)
    if smp < tot_s * 0.004 and ins < tot_i * 0.004:
        continue
    ops = {}
    for r in chunk:
        t = r[ix["Source"]].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[ix["Instructions Executed"]])
    top = sorted(ops.items(), key=lambda x: -x[1])[:5]
    stalls = {}
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stalls[h[6:]] = sum(int(r[ix[h]] or 0) for r in chunk)
    st = sorted(stalls.items(), key=lambda x: -x[1])[:4]
    print("pc %6x  samples %5.1f%%  instr %5.1f%%  top %s  stalls %s" % (
        int(chunk[0][ix["Address"]], 16) - base, 100.0 * smp / tot_s, 100.0 * ins / tot_i,
        " ".join("%s:%d%%" % (o, 100 * c // max(ins, 1)) for o, c in top), " ".join("%s:%d" % x for x in st)))
