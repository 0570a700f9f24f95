"""Radiative-effect decomposition: the host-side mirror of the reference's
RadDecomp_functions.py against golden vectors produced by importing the reference itself
(tests/golden/make_golden_raddecomp.py), bit for bit; plus the statistics columns of calc_standard_stats from the
{sum, sum of squares, count, min, max} the device reduction delivers."""
import os

import numpy as np
import pytest

from conftest import ROOT

from wrfchem_arc_interactions_b200 import decomposition as D

GOLD = os.path.join(ROOT, "tests", "golden", "raddecomp_golden.npz")
FUNCS = ("calc_Delta_S", "calc_Delta_L", "calc_SW_DIRECT", "calc_SW_INDIRECT", "calc_SW_SEMIDIRECT", "calc_LW_INDIRECT", "calc_LW_SEMIDIRECT")


def load_case(z, case):
    dicts = {}
    for name in ("BASE", "ALT"):
        d = {}
        for key in z.files:
            parts = key.split("/")
            if parts[0] == "in" and int(parts[1]) == case and parts[2] == name:
                d.setdefault(parts[3], {})[parts[4]] = z[key]
        dicts[name] = d
    return dicts["BASE"], dicts["ALT"]


@pytest.mark.parametrize("case", [0, 1, 2])
@pytest.mark.parametrize("error_type", D.ERROR_TYPES)
def test_decomposition_matches_the_reference_bit_for_bit(case, error_type):
    z = np.load(GOLD)
    B, A = load_case(z, case)
    assert set(B) == {v + s for v in ("SWUPT", "LWUPT", "LWUPTC", "SWUPTCLN") for s in ("", "_nA")}
    for fn in FUNCS:
        eff, err = getattr(D, fn)(B, A, error_type)
        assert np.array_equal(eff, z["out/%d/%s/%s/effect" % (case, error_type, fn)]), fn
        assert np.array_equal(err, z["out/%d/%s/%s/error" % (case, error_type, fn)]), fn


def test_reference_lw_indirect_quirk_and_fixed_formula():
    """The reference reads LWUPTC_nA for both operands (RadDecomp_functions.py:218-221): effect identically zero.  fixed=True
    evaluates the formula of its own comment, and then the terms close: (cloud LW effect change) = indirect + semi-direct."""
    z = np.load(GOLD)
    B, A = load_case(z, 1)
    eff, _ = D.calc_LW_INDIRECT(B, A, "standard_error")
    assert np.all(eff == 0.0)
    ind, _ = D.calc_LW_INDIRECT(B, A, "standard_error", fixed=True)
    assert np.any(ind != 0.0)
    semi, _ = D.calc_LW_SEMIDIRECT(B, A, "standard_error")
    total = (B["LWUPTC"]["avg"] - B["LWUPT"]["avg"]) - (A["LWUPTC"]["avg"] - A["LWUPT"]["avg"])
    np.testing.assert_allclose(ind + semi, total, rtol=0, atol=1e-10)


def test_sw_terms_close():
    """Delta_S = SW_DIRECT + SW_SEMIDIRECT + (SWUPT_ALT_nA - SWUPT_BASE_nA)  (the derivation at RadDecomp_functions.py:181-188)."""
    z = np.load(GOLD)
    B, A = load_case(z, 2)
    ds, _ = D.calc_Delta_S(B, A, "standard_error")
    direct, _ = D.calc_SW_DIRECT(B, A, "standard_error")
    semi, _ = D.calc_SW_SEMIDIRECT(B, A, "standard_error")
    np.testing.assert_allclose(direct + semi + (A["SWUPT_nA"]["avg"] - B["SWUPT_nA"]["avg"]), ds, rtol=0, atol=1e-10)


def test_stats_from_sums_are_the_ncl_columns():
    """avg, stddev (N-1), min, max, SE = stddev/sqrt(N), N of calc_standard_stats (misc_stats_library.ncl:432-446)."""
    rng = np.random.default_rng(5)
    fields = [rng.normal(250.0, 30.0, (40, 30)).astype(np.float32) for _ in range(3)]
    sums = np.array([[f.astype(np.float64).sum(), (f.astype(np.float64) ** 2).sum(), f.size, f.min(), f.max()] for f in fields])
    st = D.stats_from_sums(sums, names=("a", "b", "c"))
    for nm, f in zip(("a", "b", "c"), fields):
        f64 = f.astype(np.float64)
        assert st[nm]["N"] == f.size
        np.testing.assert_allclose(st[nm]["avg"], f64.mean(), rtol=1e-13)
        np.testing.assert_allclose(st[nm]["stddev"], f64.std(ddof=1), rtol=1e-9)
        np.testing.assert_allclose(st[nm]["standard_error"], f64.std(ddof=1) / np.sqrt(f.size), rtol=1e-9)
        assert st[nm]["min"] == f.min() and st[nm]["max"] == f.max()
    # a scenario dict built from such statistics feeds the decomposition directly
    B = {"SWUPT": st["a"]}; A = {"SWUPT": st["b"]}
    eff, err = D.calc_Delta_S(B, A, "standard_error")
    assert eff == st["b"]["avg"] - st["a"]["avg"] and err > 0
