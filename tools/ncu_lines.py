#!/usr/bin/env python
"""Correlate an `ncu --page source --csv` SASS dump with source lines (nvdisasm -gi of the matching object).

usage: ncu_lines.py <source.csv> <object.o> <kernel mangled-name substring> [topN]
Prints warp instructions executed and stall samples per source line (inlined frames attributed to the innermost line).
"""
import csv, re, subprocess, sys
csvf, obj, kern = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(csvf)))
hdr = rows[1]; data = rows[2:]
for _i, _r in enumerate(data):
    if _r and _r[0] == 'Kernel Name': data = data[:_i]; break
data = [r for r in data if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
import tempfile, os, glob
if obj.endswith(".cubin"):
    cub = obj
else:
    td = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, capture_output=True)
    cub = glob.glob(td + "/*.cubin")[0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cub], capture_output=True, text=True).stdout.splitlines()
# find the function section
lines = []   # per instruction: (file, line)
cur = None; infn = False; fresh = True
for l in dis:
    if re.match(r"\s*\.section\s+", l):
        infn = (".text." in l) and (kern in l)
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        fr = (m.group(1).split("/")[-1], int(m.group(2)))
        # a run of //## lines lists the inline chain innermost first; keep the innermost frame that is in a .cu file
        if fresh: cur = fr; fresh = False
        elif not cur[0].endswith(".cu") and fr[0].endswith(".cu"): cur = fr
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur); fresh = True
print("sass instrs: csv %d, disasm %d" % (len(data), len(lines)))
n = min(len(data), len(lines))
agg = {}
tot_i = tot_s = 0
for k in range(n):
    r = data[k]
    ins = int(r[ix["Instructions Executed"]] or 0); smp = int(r[ix["# Samples"]] or 0)
    lsb = int(r[ix["stall_long_sb"]] or 0)
    a = agg.setdefault(lines[k], [0, 0, 0]); a[0] += ins; a[1] += smp; a[2] += lsb
    tot_i += ins; tot_s += smp
print("%-22s %8s %8s %8s" % ("file:line", "inst%", "samples%", "long_sb%"))
for key, (i, s, l) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print("%-22s %8.2f %8.2f %8.2f" % ("%s:%d" % key if key else "?", 100.0 * i / tot_i, 100.0 * s / tot_s, 100.0 * l / tot_s))
if len(sys.argv) > 5:
    # ranges "name:lo-hi,name:lo-hi" on the .cu file
    print("\nby range:")
    for spec in sys.argv[5].split(","):
        name, r = spec.split(":"); lo, hi = map(int, r.split("-"))
        i = sum(v[0] for k, v in agg.items() if k and k[0].endswith(".cu") and lo <= k[1] <= hi)
        s = sum(v[1] for k, v in agg.items() if k and k[0].endswith(".cu") and lo <= k[1] <= hi)
        print("  %-14s inst %6.2f%%  samples %6.2f%%" % (name, 100.0 * i / tot_i, 100.0 * s / tot_s))
    i = sum(v[0] for k, v in agg.items() if not (k and k[0].endswith(".cu")))
    s = sum(v[1] for k, v in agg.items() if not (k and k[0].endswith(".cu")))
    print("  %-14s inst %6.2f%%  samples %6.2f%%" % ("other files", 100.0 * i / tot_i, 100.0 * s / tot_s))
