#!/usr/bin/env python
"""Writes the tracked parity report (profiles/r2_parity.md): for every BASELINE configuration the CUDA path against the
oracle (C++ restatement of the v3.9.1 Fortran, all host threads) - columns compared, columns out of tolerance, worst
deviations.  Run on the GPU box: python tools/parity_report.py > gpurun_out/r2_parity.md"""
import os, sys, tempfile, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from wrfchem_arc_interactions_b200 import ktables, radiation as R
import oracle as O
from conftest import run_pair, init
from parity_cases import FULL_CASES, run_case

ktab = ktables.write_files(tempfile.mkdtemp())
lib, orc = R.lib(), O.oracle_mt(0)
print("# r2 parity report: CUDA path vs oracle, one tolerance for every column\n")
print("Tolerance: every flux output within 1e-4 relative or 0.01 W/m2, heating rates within 1e-4 or 0.01 K/day (BASELINE.json north_star).")
print("Oracle: `oracle/` C++ restatement of the v3.9.1 Fortran on %d host threads; parity is unpinned against a gfortran build (no Fortran compiler, no reference fixtures)." % orc.nthreads)
print("Synthetic k-distribution tables in the real record layout (the real RRTMG_*_DATA files are not in the reference).\n")
print("| config | what | spectrum | columns compared | out of tolerance | worst abs. deviation (W/m2) | worst rel. deviation of cells > 0.01 W/m2 | worst heating-rate deviation (K/day) | output cells not bit-exact |")
print("|---|---|---|---|---|---|---|---|---|")
rows = []
for case in FULL_CASES:
    res = run_case(case, lib, orc, ktab, run_pair, init)
    for which in ("sw", "lw"):
        r = res[which]
        n = "%d (%d sunlit)" % (r["columns"], r["sunlit"]) if which == "sw" else str(r["columns"])
        print("| %s | %s | %s | %s | %d | %.2e | %s | %.2e | %d |" % (case[0], case[5], which.upper(), n, r["out_of_tolerance"], r["worst_abs"],
                                                               ("%.2e" % r["worst_rel"]) if r["worst_rel"] > 0 else "none above 0.01", r["worst_hr"],
                                                               r["cells_not_bit_exact"]))
        rows.append(dict(config=case[0], spectrum=which, **r))
    sys.stdout.flush()
print("\n```json\n" + json.dumps(rows) + "\n```")
