#!/usr/bin/env python
"""Dynamic opcode mix of one kernel from an `ncu --page source --csv` dump.  usage: ncu_opmix.py <source.csv> [topN]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1]))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
op = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) != len(hdr) or not r[ix['Instructions Executed']].isdigit():
        continue
    src = r[ix['Source']].split(); n = int(r[ix['Instructions Executed']])
    o = (src[1] if src[0].startswith('@') else src[0]).split('.')[0]
    op[o] += n; tot += n
print("warp instructions", tot)
for k, v in op.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print("%-10s %5.1f %%" % (k, 100.0 * v / tot))
