#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts + a bench JSON line into the tracked summaries under profiles/.
usage: tools/profile_summary.py <tag> <launches.csv> <full.ncu-rep> <bench.json> [more.ncu-rep ...]"""
import collections, csv, json, os, subprocess, sys
tag, launches, rep, bench = sys.argv[1:5]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = open(os.path.join(ROOT, "profiles", "%s_summary.md" % tag), "w")
W = lambda s="": out.write(s + "\n")
W("# %s — measured on B200 (sm_100a), config C2 (425x300 columns x 50 levels), see DESIGN.md section 6\n" % tag)
b = json.loads(open(bench).read().strip().splitlines()[-1])
W("## bench.py line\n")
W("```json\n" + json.dumps(b, indent=1) + "\n```\n")
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg, cnt = collections.OrderedDict(), {}
for r in data:
    if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
        continue
    name = r[kn].split("(")[0].replace("void ", "").replace("arc::", "")
    agg[name] = agg.get(name, 0) + float(r[mv].replace(",", "")); cnt[name] = cnt.get(name, 0) + 1
tot = sum(agg.values())
W("## ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`, raw CSV: profiles/%s_launches.csv)\n" % tag)
W("Per-launch times under ncu are cold-cache and serialised: compare shares.\n")
W("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda x: -x[1]):
    if v / tot > 0.0005:
        W("| `%s` | %d | %.2f | %.1f %% |" % (k, cnt[k], v / 1e6, 100 * v / tot))
W()
km = b.get("kernel_ms_per_step", {})
if km:
    t = sum(km.values())
    W("Live CUDA-event shares inside bench.py (same command without ncu): " + ", ".join("%s %.1f %%" % (k, 100 * v / t) for k, v in km.items()) + "\n")
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
reps = [rep] + sys.argv[5:]
W("## ncu --set full (launches of each hot kernel, tile 256x128x50 = 32768 columns; reports: %s under gpurun_out/, not tracked)\n" % ", ".join(os.path.basename(x) for x in reps))
det = ""
for rp in reps:
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u = rr[0], rr[1]
    for r in rr[2:]:
        W("### `%s`\n" % r[h.index("Kernel Name")])
        W("| metric | value | unit |\n|---|---|---|")
        for w in want:
            if w in h:
                W("| %s | %s | %s |" % (w, r[h.index(w)], u[h.index(w)]))
        W()
    det += subprocess.run(["ncu", "-i", rp, "--page", "details"], capture_output=True, text=True).stdout
keep = [l for l in det.splitlines() if any(s in l for s in ("k_sw_solve", "k_lw_solve", "k_sw_sweep", "k_lw_sweep", "k_sw_reduce", "k_lw_reduce", "Stall", "stalled", "Issue Slots Busy", "No Eligible", "Warp Cycles Per Issued", "Achieved Occupancy", "L2 Hit Rate", "DRAM Throughput"))]
W("## ncu details excerpts\n\n```\n" + "\n".join(keep) + "\n```")
out.close()
import shutil
shutil.copy(launches, os.path.join(ROOT, "profiles", "%s_launches.csv" % tag))
print("wrote profiles/%s_summary.md" % tag)
