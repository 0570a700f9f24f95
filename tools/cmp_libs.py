#!/usr/bin/env python
"""Bitwise comparison of two builds of libarcrad.so on the same synthetic tile.
    ARC_RAD_LIB=<a.so> python tools/cmp_libs.py dump a.npz [ni nj nk]
    ARC_RAD_LIB=<b.so> python tools/cmp_libs.py dump b.npz [ni nj nk]
    python tools/cmp_libs.py cmp a.npz b.npz"""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if sys.argv[1] == "cmp":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = 0
    for k in a.files:
        same = np.array_equal(a[k], b[k], equal_nan=True)
        if not same:
            d = np.abs(a[k].astype(np.float64) - b[k]); r = d / np.maximum(np.abs(a[k]), 1e-30)
            print("DIFF %-12s max abs %.3e  max rel %.3e  cells %d / %d" % (k, np.nanmax(d), np.nanmax(r[d > 0]) if (d > 0).any() else 0, int((d > 0).sum()), d.size))
            bad += 1
    print("identical" if not bad else "%d arrays differ" % bad)
    sys.exit(0)
from wrfchem_arc_interactions_b200 import synth, ktables, radiation as R
ni, nj, nk = [int(x) for x in (sys.argv[3:6] if len(sys.argv) > 5 else (96, 64, 50))]
d = tempfile.mkdtemp(); psw, plw = ktables.write_files(d)
dom = synth.make_domain(ni, nj, nk)
lib = R.lib(); lib.init(dom["p_top"], dom["dims"]["kme"], psw, plw)
flags = R.common_flags(dom)
o_sw, o_lw = R.alloc_outputs(dom, "sw"), R.alloc_outputs(dom, "lw")
lib.RRTMG_LWRAD(dom["dims"], **R.lw_kwargs(dom, o_lw, **flags))
lib.RRTMG_SWRAD(dom["dims"], **R.sw_kwargs(dom, o_sw, **flags))
np.savez(sys.argv[2], **{k: v for k, v in o_sw.items()}, **{k: v for k, v in o_lw.items()})
print("wrote", sys.argv[2])
