#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: stall totals by reason and the hottest SASS instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
for _i, _r in enumerate(data):
    if _r and _r[0] == 'Kernel Name': data = data[:_i]; break
data = [r for r in data if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in reasons}
for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
    print("  %-24s %6.2f%%" % (h, 100.0 * v / max(tot, 1)))
ninst = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print("warp instructions", ninst)
# opcode histogram
ops = {}
for r in data:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"): op = r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + int(r[ix["Instructions Executed"]] or 0)
print("opcode mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / ninst) for k, v in sorted(ops.items(), key=lambda x: -x[1])[:16]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:n]:
    top = sorted(((int(r[ix[h]] or 0), h) for h in reasons), reverse=True)[0]
    print("%5.2f%%  %-70s %s" % (100.0 * int(r[ix["# Samples"]]) / tot, r[ix["Source"]].strip()[:70], top[1]))
