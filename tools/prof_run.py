#!/usr/bin/env python
"""Developer script for ncu: device-resident SW+LW steps on a synthetic tile (no oracle, no host copies)."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from wrfchem_arc_interactions_b200 import synth, ktables, radiation as R, abi
ni, nj, nk = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (128, 128, 50))]
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
d = tempfile.mkdtemp(); psw, plw = ktables.write_files(d)
dom = synth.make_domain(ni, nj, nk)
lib = R.lib(); lib.init(dom["p_top"], dom["dims"]["kme"], psw, plw, device=0)
dev = torch.device("cuda", 0)
ddom = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) and v.ndim >= 2 else v) for k, v in dom.items()}
like = ddom["xcoszen"]
o_sw, o_lw = R.alloc_outputs(dom, "sw", like=like), R.alloc_outputs(dom, "lw", like=like)
flags = R.common_flags(dom)
dims = abi.make_dims(dom["dims"])
for _ in range(steps):
    lib.RRTMG_LWRAD(dims, **R.lw_kwargs(ddom, o_lw, **flags))
    lib.RRTMG_SWRAD(dims, **R.sw_kwargs(ddom, o_sw, **flags))
torch.cuda.synchronize()
for n in ("sw_mcica", "sw_prep", "sw_solve", "sw_sweep", "sw_reduce", "lw_mcica", "lw_prep", "lw_solve", "lw_sweep", "lw_reduce"):
    print(n, lib.lib.arc_rad_last_kernel_ms(n.encode()))
print("swupt mean", float(o_sw["swupt"].mean()), "olr mean", float(o_lw["olr"].mean()))
