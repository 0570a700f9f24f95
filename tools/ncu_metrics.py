#!/usr/bin/env python
"""profiles/r2_ncu_metrics.json from `ncu --set full` reports: per hot kernel the DRAM bytes per column, issue-slot
utilisation, executed thread-instructions per (column, g-point, layer) and lane utilisation of ONE profiled launch.
bench.py copies `dram_bytes_per_column` / `issue_active_pct` into its roofline block (with `_source`), so the numbers in
the bench line are tied to a tracked capture and the commit it was taken at instead of being constants in the script.

usage: tools/ncu_metrics.py <ncols of the profiled tile> <sunlit columns per SW launch> <nk> <nlay_lw> <rep> [<rep> ...]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ncols, nsun, nk, nlay_lw = [int(x) for x in sys.argv[1:5]]
out = {}
for rp in sys.argv[5:]:
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h = rr[0]
    for r in rr[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("arc::", "").split("<")[0]
        if name in out:
            continue
        def val(k):
            return float(r[h.index(k)].replace(",", "")) if k in h else None
        unit = {k: rr[1][h.index(k)] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in h}
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = val("dram__bytes_read.sum") * scale.get(unit.get("dram__bytes_read.sum", "byte"), 1.0)
        wr = val("dram__bytes_write.sum") * scale.get(unit.get("dram__bytes_write.sum", "byte"), 1.0)
        sw = name.startswith("k_sw")
        aer = name.startswith("k_aer")
        cols = nsun if sw else (int(os.environ.get("AER_COLS", "8192")) if aer else ncols)
        cells = cols * (112.0 * (nk + 1) if sw else (8.0 * 20 * nk if aer else 140.0 * nlay_lw))     # aerosol: per (section, wavelength, level)
        winst = val("smsp__inst_executed.sum")
        lanes = val("smsp__thread_inst_executed_per_inst_executed.ratio")
        out[name] = {"columns_in_launch": cols, "duration_ms_under_ncu": val("gpu__time_duration.sum") / (1e6 if rr[1][h.index("gpu__time_duration.sum")] == "ns" else 1e3 if rr[1][h.index("gpu__time_duration.sum")] == "us" else 1.0),
                     "dram_bytes_per_column": (rd + wr) / cols, "dram_read_bytes": rd, "dram_write_bytes": wr,
                     "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
                     "registers": val("launch__registers_per_thread"),
                     "thread_instructions_per_cell": winst * lanes / cells if winst and lanes else None,
                     "issued_warp_instructions_x32_per_cell": winst * 32.0 / cells if winst else None,
                     "active_lanes_of_32": lanes, "report": os.path.basename(rp)}
try:
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
except Exception:
    commit = None
out["_cell"] = "cell = (column, g-point, layer) for the RRTMG kernels, (column, section, wavelength, level) for the aerosol kernels; k_sw_sweep: one of the 21 sweep-group launches"
out["_source"] = "profiles/r2_ncu_metrics.json (ncu --set full --clock-control none, tile %d columns x %d levels, at commit %s)" % (ncols, nk, commit)
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_ncu_metrics.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
