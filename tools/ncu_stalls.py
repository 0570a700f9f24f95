#!/usr/bin/env python
"""Stall-reason breakdown per source-line range.  usage: ncu_stalls.py <source.csv> <object.o> <kernel substring> name:lo-hi,..."""
import csv, re, subprocess, sys, tempfile, os, glob
csvf, obj, kern, spec = sys.argv[1:5]
rows = list(csv.reader(open(csvf))); hdr = rows[1]; data = rows[2:]
for _i, _r in enumerate(data):
    if _r and _r[0] == 'Kernel Name': data = data[:_i]; break
data = [r for r in data if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
td = tempfile.mkdtemp(); subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, capture_output=True)
cub = glob.glob(td + "/*.cubin")[0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cub], capture_output=True, text=True).stdout.splitlines()
lines = []; cur = None; infn = False; fresh = True
for l in dis:
    if re.match(r"\s*\.section\s+", l): infn = (".text." in l) and (kern in l); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        fr = (m.group(1).split("/")[-1], int(m.group(2)))
        if fresh: cur = fr; fresh = False
        elif not cur[0].endswith(".cu") and fr[0].endswith(".cu"): cur = fr
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l): lines.append(cur); fresh = True
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
n = min(len(data), len(lines))
for sp in spec.split(","):
    name, r = sp.split(":"); lo, hi = map(int, r.split("-"))
    tot = {h: 0 for h in st}; ins = 0
    for k in range(n):
        if lines[k] and lines[k][0].endswith(".cu") and lo <= lines[k][1] <= hi:
            ins += int(data[k][ix["Instructions Executed"]] or 0)
            for h in st: tot[h] += int(data[k][ix[h]] or 0)
    T = sum(tot.values()) or 1
    top = sorted(tot.items(), key=lambda kv: -kv[1])[:6]
    print("%-8s inst %10d samples %8d : " % (name, ins, T) + "  ".join("%s %.0f%%" % (h[6:], 100.0 * v / T) for h, v in top))
