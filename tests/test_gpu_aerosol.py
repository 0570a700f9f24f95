"""Aerosol optical-property stage on the GPU (k_aer_prep + k_aer_mie through arc_aer_optics) against the oracle's independent
double-precision restatement on identical synthetic aerosol fields, and its Chebyshev tables against direct Mie theory.
PARITY UNPINNED w.r.t. the Fortran (module_optical_averaging.F is not in the reference repository): self-consistent only."""
import ctypes as C

import numpy as np
import pytest

from conftest import init, run_pair
from wrfchem_arc_interactions_b200 import abi, radiation as R, synth

pytestmark = pytest.mark.gpu


def setup(lib, orc, ktab, dom):
    init(lib, dom, ktab)
    nr = np.zeros((9, 20), np.float32); ni = np.zeros((9, 20), np.float32)
    lib.lib.arc_aer_default_refindex.restype = None
    lib.lib.arc_aer_default_refindex.argtypes = [abi.c_fp, abi.c_fp]
    lib.lib.arc_aer_default_refindex(abi.fptr(nr), abi.fptr(ni))
    assert orc.lib.arc_oracle_aer_init(abi.fptr(nr), abi.fptr(ni)) == 0
    lib.lib.arc_aer_init.restype = C.c_int
    lib.lib.arc_aer_init.argtypes = [abi.c_fp, abi.c_fp]
    lib.check(lib.lib.arc_aer_init(None, None))
    return nr, ni


def compare(og, oo, nk, loose=1.0):
    for k in og:
        a, b = og[k][:, :nk].astype(np.float64), oo[k][:, :nk].astype(np.float64)
        assert np.all(np.isfinite(a)), k
        tol = loose * (3e-4 if k.startswith(("tau", "ext")) else 2e-4)
        scale = np.abs(b).max()
        assert np.all(np.abs(a - b) <= tol * np.abs(b) + 1e-6 * scale), "%s: max rel dev %.2e" % (k, (np.abs(a - b) / np.maximum(np.abs(b), 1e-6 * scale)).max())
        assert not og[k][:, nk:].any(), "%s: level kme written" % k


@pytest.mark.parametrize("nbin", [4, 8])
def test_sectional_vs_oracle(lib, orc, ktab, nbin):
    dom = synth.make_domain(16, 6, 30, seed=41)
    setup(lib, orc, ktab, dom)
    bins, alt, _ = synth.make_aerosol(dom, nbin=nbin)
    og, oo = R.alloc_aer_outputs(dom, ext=True), R.alloc_aer_outputs(dom, ext=True)
    lib.optical_averaging(dom["dims"], "sectional", bins, alt, dom["dz8w"], og)
    orc.optical_averaging(dom["dims"], "sectional", bins, alt, dom["dz8w"], oo)
    compare(og, oo, 30)
    assert og["tauaer400"].max() > 1e-4


def test_modal_vs_oracle(lib, orc, ktab):
    dom = synth.make_domain(16, 6, 30, seed=42)
    setup(lib, orc, ktab, dom)
    modes, alt, sig = synth.make_aerosol(dom, modal=True)
    og, oo = R.alloc_aer_outputs(dom), R.alloc_aer_outputs(dom)
    lib.optical_averaging(dom["dims"], "modal", modes, alt, dom["dz8w"], og, sigmag=sig)
    orc.optical_averaging(dom["dims"], "modal", modes, alt, dom["dz8w"], oo, sigmag=sig)
    compare(og, oo, 30, loose=10.0)      # single-precision erf differences of neighbouring section edges


def test_empty_and_ragged_sections(lib, orc, ktab):
    """Edge cases of the sectional input: whole sections empty, points without any aerosol, one populated section, a single
    species per section, and a tile whose point count is not a multiple of the 256-point block - against the oracle, and
    tau = 0, omega = 1, g = 0 where nothing is there.  (The block's counting sort sees empty, single and odd-length buckets.)"""
    dom = synth.make_domain(13, 7, 23, seed=43)                     # 2,093 points: 8 full blocks + a ragged one
    setup(lib, orc, ktab, dom)
    bins, alt, _ = synth.make_aerosol(dom, nbin=8)
    for name in bins[2]:
        bins[2][name] = np.zeros_like(bins[2][name])                 # an empty section
    for name in bins[5]:
        if name not in ("num", "so4"):
            bins[5][name] = np.zeros_like(bins[5][name])             # a single-species section
    for b in bins:
        for name in b:
            b[name][:, :, 3] = 0.0                                   # a column without aerosol
            b[name][2, 5:9, :] = 0.0                                 # a slab of empty points
    for b in bins[1:]:
        for name in b:
            b[name][:, :, 7] = 0.0                                   # a column with one populated section
    og, oo = R.alloc_aer_outputs(dom, ext=True), R.alloc_aer_outputs(dom, ext=True)
    lib.optical_averaging(dom["dims"], "sectional", bins, alt, dom["dz8w"], og)
    orc.optical_averaging(dom["dims"], "sectional", bins, alt, dom["dz8w"], oo)
    compare(og, oo, 23)
    assert not og["tauaer400"][:, :23, 3].any() and np.all(og["waer400"][:, :23, 3] == 1.0) and not og["gaer400"][:, :23, 3].any()
    assert not og["tauaerlw7"][2, 5:9, :].any() and og["tauaer400"][:, :23, 7].max() > 0


def test_tables_vs_direct_mie(lib, orc, ktab):
    """Chebyshev-interpolated efficiencies against direct Mie sums at random sizes / refractive indices inside the table."""
    dom = synth.make_domain(4, 2, 10, seed=43)
    nr, ni = setup(lib, orc, ktab, dom)
    L = lib.lib
    for fn in (L.arc_aer_table_eval, L.arc_aer_mie_direct):
        fn.restype = C.c_int
        fn.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, abi.c_fp, abi.c_fp, abi.c_fp]
    rng = np.random.default_rng(5)
    errs = []
    for _ in range(400):
        wl = int(rng.integers(0, 20))
        r = float(10 ** rng.uniform(np.log10(0.02e-4), np.log10(5e-4)))
        re = float(rng.uniform(nr[:, wl].min(), nr[:, wl].max()))
        im = float(10 ** rng.uniform(np.log10(max(ni[:, wl].min(), 1e-9)), np.log10(ni[:, wl].max())))
        a = [np.zeros(1, np.float32) for _ in range(3)]; b = [np.zeros(1, np.float32) for _ in range(3)]
        assert L.arc_aer_table_eval(wl, r, re, im, *[abi.fptr(x) for x in a]) == 0
        assert L.arc_aer_mie_direct(wl, r, re, im, *[abi.fptr(x) for x in b]) == 0
        errs.append([abs(a[0][0] - b[0][0]) / b[0][0], abs(a[1][0] - b[1][0]) / max(b[1][0], 1e-12), abs(a[2][0] - b[2][0])])
    errs = np.array(errs)
    # the 50-term fit smooths the Mie ripple of weakly absorbing spheres: median error < 1 %, 95th percentile ~7 % (ripple + 7x7 refractive-index grid)
    assert np.median(errs[:, 0]) < 0.01 and np.quantile(errs[:, 0], 0.95) < 0.12
    assert np.median(errs[:, 1]) < 0.01 and np.median(errs[:, 2]) < 0.01


def test_optics_feed_the_radiation_calls(lib, orc, ktab):
    """The stage's outputs are exactly the aerosol inputs of RRTMG_SWRAD / RRTMG_LWRAD: chain them and check that the
    aerosol dims the clear-sky surface flux and that clean-sky diagnostics differ from the full ones."""
    dom = synth.make_domain(16, 6, 40, seed=44, all_day=True)
    setup(lib, orc, ktab, dom)
    bins, alt, _ = synth.make_aerosol(dom, nbin=8)
    o = R.alloc_aer_outputs(dom)
    lib.optical_averaging(dom["dims"], "sectional", bins, alt, dom["dz8w"], o)
    d2 = dict(dom); d2.update(o)
    sw, lw = run_pair("sw", lib, d2), run_pair("lw", lib, d2)
    assert np.all(sw["swdnbc"] < sw["swdnbclnc"]) and np.all(sw["swupt"] != sw["swuptcln"])
    assert np.all(lw["lwdnb"] >= lw["lwdnbcln"] - 1e-3) and (lw["lwdnb"] != lw["lwdnbcln"]).mean() > 0.9
