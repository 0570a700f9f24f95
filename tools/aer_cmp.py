#!/usr/bin/env python
"""Developer script: run the aerosol optics stage of one library build in a subprocess-free way and dump its outputs, or
compare two dumps bit for bit.   usage: aer_cmp.py run <out.npz> [ni nj nk] | aer_cmp.py cmp a.npz b.npz"""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if sys.argv[1] == "cmp":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = 0
    for k in a.files:
        n = int((a[k].view(np.uint32) != b[k].view(np.uint32)).sum())
        if n: print(k, "cells differ", n, "max abs", float(np.abs(a[k] - b[k]).max())); bad += n
    print("aer_cmp: arrays", len(a.files), "cells differing", bad)
    sys.exit(1 if bad else 0)
import torch
from wrfchem_arc_interactions_b200 import synth, ktables, radiation as R
out = sys.argv[2]
ni, nj, nk = [int(x) for x in (sys.argv[3:6] if len(sys.argv) > 5 else (96, 37, 50))]
mode = sys.argv[6] if len(sys.argv) > 6 else "sectional"
d = tempfile.mkdtemp(); psw, plw = ktables.write_files(d)
dom = synth.make_domain(ni, nj, nk)
lib = R.lib(); lib.init(dom["p_top"], dom["dims"]["kme"], psw, plw, device=0)
dev = torch.device("cuda", 0)
modal = mode == "modal"
bins, alt, sg = synth.make_aerosol(dom, nbin=8, modal=modal)
dbins = [{k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in b.items()} for b in bins]
dalt = torch.from_numpy(alt).to(dev); ddz = torch.from_numpy(dom["dz8w"]).to(dev)
like = torch.from_numpy(dom["xcoszen"]).to(dev)
ao = R.alloc_aer_outputs(dom, like=like)
for _ in range(3):
    lib.optical_averaging(abi.make_dims(dom["dims"]) if False else dom["dims"], mode, dbins, dalt, ddz, ao, sigmag=sg if modal else None)
torch.cuda.synchronize()
print(mode, ni, nj, nk, "aer_optics ms", lib.lib.arc_rad_last_kernel_ms(b"aer_optics"))
np.savez(out, **{k: v.cpu().numpy() for k, v in ao.items()})
