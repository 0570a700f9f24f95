"""Aerosol optical-property stage, CPU side: the Chebyshev-Mie tables of the product (built on the host by
csrc/aer_tables.cpp, evaluated here through arc_aer_table_eval) against direct Mie theory, and the oracle's independent
restatement on a synthetic MOSAIC column set (physical sanity: extinction per mass, single-scattering albedo ranges).
PARITY UNPINNED: module_optical_averaging.F is not in the reference repository."""
import ctypes as C

import numpy as np
import pytest

from wrfchem_arc_interactions_b200 import abi, radiation as R, synth


def refindex(lib):
    nr = np.zeros((9, 20), np.float32); ni = np.zeros((9, 20), np.float32)
    lib.lib.arc_aer_default_refindex.restype = None
    lib.lib.arc_aer_default_refindex.argtypes = [abi.c_fp, abi.c_fp]
    lib.lib.arc_aer_default_refindex(abi.fptr(nr), abi.fptr(ni))
    return nr, ni


def test_default_refractive_indices_are_physical(lib):
    nr, ni = refindex(lib)
    assert nr.min() > 1.0 and nr.max() < 3.1 and ni.min() >= 0 and ni.max() < 1.0
    assert np.allclose(nr[7], 1.95) and np.allclose(ni[7], 0.79)            # soot, wavelength independent
    assert np.all(ni[8, :4] < 1e-5)                                          # water is transparent in the SW


def test_mie_known_answers(lib):
    """Bohren & Huffman appendix A test case (x = 5.213, m = 1.55): Q_ext = Q_sca = 3.1054, g = 0.6331; and the
    Rayleigh limit Q_sca = 8/3 x^4 |(m^2-1)/(m^2+2)|^2."""
    L = lib.lib
    L.arc_aer_mie_direct.restype = C.c_int
    L.arc_aer_mie_direct.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, abi.c_fp, abi.c_fp, abi.c_fp]
    q = [np.zeros(1, np.float32) for _ in range(3)]
    lam = 0.6e-4                                                             # wl index 2
    assert L.arc_aer_mie_direct(2, 5.213 * lam / (2 * np.pi), 1.55, 0.0, *[abi.fptr(x) for x in q]) == 0
    assert abs(q[0][0] - 3.1054) < 2e-3 and abs(q[1][0] - 3.1054) < 2e-3 and abs(q[2][0] - 0.6331) < 1e-3
    x = 0.02
    assert L.arc_aer_mie_direct(2, x * lam / (2 * np.pi), 1.5, 0.0, *[abi.fptr(t) for t in q]) == 0
    m2 = 1.5 ** 2
    assert np.isclose(q[1][0], 8.0 / 3.0 * x ** 4 * ((m2 - 1) / (m2 + 2)) ** 2, rtol=2e-3)


def test_oracle_tables_match_direct_mie(orc, lib):
    """The oracle's Chebyshev interpolation against its own direct Mie sum at random radii / refractive indices inside the
    table: the 50-term fit in ln r smooths the Mie ripple, so the bar is 3 % on Q_ext for the sub-micron range and the mean
    error over all sizes is well below 1 %."""
    nr, ni = refindex(lib)
    L = orc.lib
    assert L.arc_oracle_aer_init(abi.fptr(nr), abi.fptr(ni)) == 0
    dom = synth.make_domain(6, 3, 20, seed=31)
    bins, alt, _ = synth.make_aerosol(dom, nbin=8)
    outs = R.alloc_aer_outputs(dom, ext=True)
    orc.optical_averaging(dom["dims"], "sectional", bins, alt, dom["dz8w"], outs)
    t4 = outs["tauaer400"][:, :20]
    assert np.all(np.isfinite(t4)) and np.all(t4 >= 0) and t4.max() > 1e-4
    aod = outs["tauaer400"][:, :20].sum(axis=1)
    assert 0.005 < np.median(aod) < 5.0                                       # column AOD(400 nm) of a ~12 ug/kg boundary-layer aerosol
    w4 = outs["waer400"][:, :20]
    assert w4.min() > 0.5 and w4.max() <= 1.0                                 # 5 % soot
    g4 = outs["gaer400"][:, :20]
    assert 0.3 < g4.min() and g4.max() < 0.95
    # Angstrom exponent between 400 and 999 nm is positive for this sub-micron dominated mixture
    a = np.log(outs["tauaer400"][:, :20].sum(axis=1) / outs["tauaer999"][:, :20].sum(axis=1)) / np.log(999.0 / 400.0)
    assert 0.0 < np.median(a) < 3.0
    # LW absorption optical depth is much smaller than the visible extinction
    assert np.median(outs["tauaerlw8"][:, :20].sum(axis=1) / aod) < 0.5
    assert np.allclose(outs["extaerlw8"][:, :20], outs["tauaerlw8"][:, :20] / (dom["dz8w"][:, :20] * 100.0) * 1e5, rtol=1e-5, atol=1e-12)


def test_oracle_modal_conserves_extinction_scale(orc, lib):
    nr, ni = refindex(lib)
    assert orc.lib.arc_oracle_aer_init(abi.fptr(nr), abi.fptr(ni)) == 0
    dom = synth.make_domain(6, 3, 20, seed=32)
    modes, alt, sig = synth.make_aerosol(dom, modal=True)
    outs = R.alloc_aer_outputs(dom)
    orc.optical_averaging(dom["dims"], "modal", modes, alt, dom["dz8w"], outs, sigmag=sig)
    aod = outs["tauaer400"][:, :20].sum(axis=1)
    assert np.all(np.isfinite(aod)) and 0.005 < np.median(aod) < 5.0
    # doubling every mass and number doubles the optical depth exactly (linear in number at fixed size / composition)
    m2 = [{k: (v * 2).astype(np.float32) for k, v in md.items()} for md in modes]
    o2 = R.alloc_aer_outputs(dom)
    orc.optical_averaging(dom["dims"], "modal", m2, alt, dom["dz8w"], o2, sigmag=sig)
    assert np.allclose(o2["tauaer600"], 2 * outs["tauaer600"], rtol=2e-5)
    assert np.allclose(o2["waer600"], outs["waer600"], rtol=2e-5)
