#!/usr/bin/env python
"""Developer script for ncu: aerosol optics stage device-resident."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from wrfchem_arc_interactions_b200 import synth, ktables, radiation as R, abi
ni, nj, nk = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (128, 64, 50))]
d = tempfile.mkdtemp(); psw, plw = ktables.write_files(d)
dom = synth.make_domain(ni, nj, nk)
lib = R.lib(); lib.init(dom["p_top"], dom["dims"]["kme"], psw, plw, device=0)
dev = torch.device("cuda", 0)
bins, alt, _ = synth.make_aerosol(dom, nbin=8)
dbins = [{k: torch.from_numpy(v).to(dev) for k, v in b.items()} for b in bins]
dalt = torch.from_numpy(alt).to(dev); ddz = torch.from_numpy(dom["dz8w"]).to(dev)
like = torch.from_numpy(dom["xcoszen"]).to(dev)
ao = R.alloc_aer_outputs(dom, like=like)
for _ in range(3):
    lib.optical_averaging(dom["dims"], "sectional", dbins, dalt, ddz, ao)
torch.cuda.synchronize()
print("aer_optics ms", lib.lib.arc_rad_last_kernel_ms(b"aer_optics"), "aod", float(ao["tauaer400"].sum(dim=1).median()))
