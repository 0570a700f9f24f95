#!/usr/bin/env python
"""Generate tests/golden/oracle_c1_slice.npz from the oracle (run in the build container; the reference itself ships no
golden vectors and cannot be compiled here -- no Fortran compiler -- so these pin the oracle against regressions only)."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from wrfchem_arc_interactions_b200 import synth, ktables, radiation as R
import oracle as O
ni, nj, nk, seed = 8, 4, 40, 2012
psw, plw = ktables.write_files(tempfile.mkdtemp())
dom = synth.make_domain(ni, nj, nk, seed=seed)
orc = O.oracle(); orc.init(dom["p_top"], dom["dims"]["kme"], psw, plw)
flags = R.common_flags(dom)
sw, lw = R.alloc_outputs(dom, "sw"), R.alloc_outputs(dom, "lw")
orc.RRTMG_SWRAD(dom["dims"], **R.sw_kwargs(dom, sw, **flags)); orc.RRTMG_LWRAD(dom["dims"], **R.lw_kwargs(dom, lw, **flags))
out = dict(ni=ni, nj=nj, nk=nk, seed=seed)
for k in ("swupt", "swuptc", "swuptcln", "swdnb", "swdnbc", "swdnbcln", "gsw", "swcf", "swddir", "swddif", "rthratensw", "swupflx", "swdnflxcln"):
    out["sw_" + k] = sw[k]
for k in ("lwupt", "lwuptc", "lwuptcln", "lwdnb", "lwdnbc", "lwdnbcln", "glw", "olr", "lwcf", "rthratenlw", "lwupflx", "lwdnflxcln"):
    out["lw_" + k] = lw[k]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_c1_slice.npz"), **out)
print("written", {k: getattr(v, "shape", v) for k, v in out.items()})
