#!/usr/bin/env python
"""Developer script: run the CUDA library and the oracle on the same synthetic tile and print error statistics."""
import os, sys, time, tempfile, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import wrfchem_arc_interactions_b200 as pkg
from wrfchem_arc_interactions_b200 import synth, ktables, radiation as R, abi
import oracle as O

ni, nj, nk = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (32, 32, 40))]
kw = dict(cloudy_frac=0.4)
if len(sys.argv) > 4 and sys.argv[4] == "re":
    kw = dict(cloudy_frac=1.0, with_re=True)
d = tempfile.mkdtemp()
psw, plw = ktables.write_files(d)
dom = synth.make_domain(ni, nj, nk, **kw)
lib = R.lib(); orc = O.oracle()
lib.init(dom["p_top"], dom["dims"]["kme"], psw, plw)
orc.init(dom["p_top"], dom["dims"]["kme"], psw, plw)
flags = R.common_flags(dom)
ncol = ni * nj

def stats(name, a, b, mask=None):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    if mask is not None: a = a[mask]; b = b[mask]
    if a.size == 0: return
    err = np.abs(a - b); rel = err / np.maximum(np.abs(b), 1e-30)
    ok = (err <= 1e-4 * np.abs(b)) | (err <= 0.01)
    print("%-12s maxabs %.3e maxrel(where abs>0.01) %.3e  bad %d / %d  ref range [%.4g, %.4g]" % (
        name, err.max(), (rel * (err > 0.01)).max(), (~ok).sum(), a.size, b.min(), b.max()))

for which in ("sw", "lw"):
    nlay = nk + 1 if which == "sw" else lib.lw_nlayers()
    ng = 112 if which == "sw" else 140
    og = R.alloc_outputs(dom, which); oo = R.alloc_outputs(dom, which)
    dg, ag = abi.alloc_debug(ncol, nlay, ng); do, ao = abi.alloc_debug(ncol, nlay, ng)
    kwf = R.sw_kwargs if which == "sw" else R.lw_kwargs
    fn_g = lib.RRTMG_SWRAD if which == "sw" else lib.RRTMG_LWRAD
    fn_o = orc.RRTMG_SWRAD if which == "sw" else orc.RRTMG_LWRAD
    t = time.time(); fn_g(dom["dims"], debug=dg, **kwf(dom, og, **flags)); tg = time.time() - t
    t = time.time(); fn_o(dom["dims"], debug=do, **kwf(dom, oo, **flags)); to = time.time() - t
    print("==== %s  gpu call %.3fs  oracle %.3fs" % (which, tg, to))
    sun = ao["laytrop"] >= 0
    for k in ("laytrop", "jp", "jt", "jt1", "indfor", "indself", "indminor", "cldmask"):
        if which == "sw" and k == "indminor": continue
        m = sun if ag[k].ndim == 1 else sun.reshape((-1,) + (1,) * (ag[k].ndim - 1)) & np.ones(ag[k].shape, bool)
        print("%-10s mismatches %d / %d" % (k, int((ag[k][m] != ao[k][m]).sum()), int(m.sum())))
    for k in ("fac00", "fac01", "fac10", "fac11", "taug", "taur", "sfluxzen", "taucmc", "hr"):
        m = sun.reshape((-1,) + (1,) * (ag[k].ndim - 1)) & np.ones(ag[k].shape, bool)
        a_, b_ = ag[k][m].astype(np.float64), ao[k][m].astype(np.float64)
        err = np.abs(a_ - b_); rel = err / np.maximum(np.abs(b_), 1e-30)
        sel = np.abs(b_) > 1e-12 * max(np.abs(b_).max() if b_.size else 0, 1e-30)
        nbit = int((ag[k][m].view(np.uint32) != ao[k][m].view(np.uint32)).sum())
        print("%-10s maxabs %.3e  maxrel %.3e  (ref absmax %.4g)  bitwise different %d / %d" % (k, err.max() if err.size else 0, rel[sel].max() if sel.any() else 0, np.abs(b_).max() if b_.size else 0, nbit, a_.size))
    for k in og:
        stats(k, og[k], oo[k])
    if which == "sw":
        cond = ao["sw_cond"]
        worst = np.zeros(ncol)
        for k in ("swupt", "swuptc", "swuptcln", "swdnb", "swdnbc", "swdnbcln", "swupb", "gsw"):
            e = np.abs(og[k].astype(np.float64) - oo[k]).ravel(); r = np.abs(oo[k]).ravel()
            worst = np.maximum(worst, np.where(e > 0.01, e / np.maximum(r, 1e-9), 0.0))
        hre = np.abs(ag["hr"].astype(np.float64) - ao["hr"]).max(axis=1)
        print("cond quantiles (sunlit)", np.quantile(cond[sun], [0, 0.001, 0.01, 0.05, 0.5]))
        for thr in (1e-4, 3e-4, 1e-3, 3e-3, 1e-2):
            sel = sun & (cond >= thr)
            print("cond >= %g: kept %d of %d sunlit; max flux rel err %.3e; max |dhr| %.3e K/day" % (thr, sel.sum(), sun.sum(), worst[sel].max(), hre[sel].max()))
        bad = np.argsort(-worst)[:8]
        for w_ in bad: print("  col %d relerr %.2e dhr %.2e cond %.2e" % (w_, worst[w_], hre[w_], cond[w_]))
    key = "swupt" if which == "sw" else "lwupt"
    e = np.abs(og[key] - oo[key]).ravel(); w = int(e.argmax())
    print("worst column", w, "err", e[w], "coszen", dom["xcoszen"].ravel()[w], "albedo", dom["albedo"].ravel()[w],
          "cloudy layers", int((dom["cldfra3d"].reshape(nj, nk + 1, ni)[w // ni, :, w % ni] > 0).sum()),
          "aod400", float(dom["tauaer400"].reshape(nj, nk + 1, ni)[w // ni, :, w % ni].sum()))
    hrg, hro = ag["hr"][w], ao["hr"][w]
    print("hr gpu", np.array2string(hrg, precision=3, max_line_width=200)); print("hr orc", np.array2string(hro, precision=3, max_line_width=200))
# device logf / expf / powf against the C library
import ctypes as C
L, Ol = lib.lib, orc.lib
L.arc_rad_selftest_libm.argtypes = [C.c_int, abi.c_fp, abi.c_fp, C.c_int, abi.c_fp, C.c_int]
Ol.arc_oracle_libm.argtypes = [C.c_int, abi.c_fp, abi.c_fp, C.c_int, abi.c_fp]
rng = np.random.default_rng(3)
for which, x, y in ((0, np.exp(rng.uniform(np.log(1e-3), np.log(1200.), 1 << 22)), None), (1, rng.uniform(-87, 87, 1 << 22), None),
                    (2, np.exp(rng.uniform(-10, 10, 1 << 22)), rng.uniform(-4, 4, 1 << 22))):
    x = x.astype(np.float32); y = (y if y is not None else x).astype(np.float32)
    a = np.empty_like(x); b = np.empty_like(x)
    assert L.arc_rad_selftest_libm(which, abi.fptr(x), abi.fptr(y), x.size, abi.fptr(a), 1) == 0
    Ol.arc_oracle_libm(which, abi.fptr(x), abi.fptr(y), x.size, abi.fptr(b))
    print("libm device fn %d: %d of %d differ from the C library" % (which, int((a.view(np.uint32) != b.view(np.uint32)).sum()), x.size))
print("launches", lib.lib.arc_rad_launch_count())
for n in ("sw_mcica", "sw_prep", "sw_solve", "sw_sweep", "sw_reduce", "lw_mcica", "lw_prep", "lw_solve", "lw_sweep", "lw_reduce"):
    print(n, lib.lib.arc_rad_last_kernel_ms(n.encode()))
