"""Ghan (2012) / Archer-Nicholls (2016) decomposition of the TOA radiative effect of an emission source into direct,
semi-direct and indirect parts, from the domain statistics of four scenarios.

Host-side mirror of the reference's post-processing interface
(analysis_scripts/RadDecomp_analysis_package/RadDecomp_functions.py:119-267): same function names, same arguments
(`BASE_dict`, `ALT_dict`, `error_type`), same return value `(effect, error)`.  A "dict" maps a variable name
('SWUPT', 'SWUPTCLN', 'LWUPT', 'LWUPTC' and the same with the suffix '_nA' for the run without aerosol-radiation
interaction) to anything indexable by 'avg' and by the error column name (a pandas DataFrame in the reference; a plain
dict of numpy arrays or the output of `stats_from_sums` here).  Errors add in quadrature.

The numbers come from the device: `arc_rad_domain_stats` reduces every TOA field of a tile to
{sum, sum of squares, count, min, max} (all-reduced over the GPUs of a run by `partition.combine_stats`), and
`stats_from_sums` turns them into the columns calc_standard_stats writes
(analysis_scripts/NCL_extraction_package/misc_stats_library.ncl:396-461): avg, stddev (NCL `stddev`: N-1 in the
denominator), min, max, standard_error = stddev / sqrt(N), N; with Moran's I from `arc_rad_morans_i` also the
corrected_standard_error = SE * I column.

Reference quirk kept by default: calc_LW_INDIRECT (RadDecomp_functions.py:216-233) reads 'LWUPTC_nA' for BOTH the
clear-sky and the all-sky operand, so its effect is identically zero and its error counts LWUPTC_nA four times.
`calc_LW_INDIRECT(..., fixed=True)` evaluates the formula its own comment states, (LWUPTC - LWUPT)_BASE_nA -
(LWUPTC - LWUPT)_ALT_nA.
"""
from __future__ import annotations

import numpy as np

ERROR_TYPES = ("standard_error", "corrected_standard_error")


def stats_from_sums(sums, names=None, morans_i=None, percentiles=None):
    """{sum, sum of squares, count, min, max} per field (the `out[f][5]` of arc_rad_domain_stats) -> the statistics columns of
    calc_standard_stats that the decomposition uses.  `sums`: array (nfields, 5); `names`: field names (-> dict of dicts);
    `morans_i`: Moran's I per field (arc_rad_morans_i) -> also 'morans_i' and 'corrected_standard_error' = SE * I
    (misc_stats_library.ncl:447-449); `percentiles`: array (nfields, 5) = median, lower / upper quartile, 5th / 95th
    percentile (arc_rad_percentiles with perc = 50, 25, 75, 5, 95; ncl:439-445)."""
    s = np.asarray(sums, dtype=np.float64)
    n = s[:, 2]
    avg = s[:, 0] / n
    var = (s[:, 1] - n * avg * avg) / np.maximum(n - 1.0, 1.0)
    sd = np.sqrt(np.maximum(var, 0.0))
    cols = {"avg": avg, "stddev": sd, "min": s[:, 3], "max": s[:, 4], "standard_error": sd / np.sqrt(n), "N": n}
    if morans_i is not None:
        mi = np.asarray(morans_i, dtype=np.float64)
        cols["morans_i"] = mi
        cols["corrected_standard_error"] = cols["standard_error"] * mi
    if percentiles is not None:
        pc = np.asarray(percentiles, dtype=np.float64)
        for q, nm in enumerate(("median", "lower_quartile", "upper_quartile", "p05", "p95")):
            cols[nm] = pc[:, q]
    if names is None:
        return cols
    return {nm: {k: v[i] for k, v in cols.items()} for i, nm in enumerate(names)}


def _quad(*errs):
    acc = np.power(errs[0], 2)
    for e in errs[1:]:
        acc = acc + np.power(e, 2)
    return np.sqrt(acc)


# ---- net effects at TOA (RadDecomp_functions.py:119-143) ---------------------------------------------------------------
def calc_Delta_S(BASE_dict, ALT_dict, error_type):
    b, a = BASE_dict["SWUPT"], ALT_dict["SWUPT"]
    return a["avg"] - b["avg"], _quad(a[error_type], b[error_type])


def calc_Delta_L(BASE_dict, ALT_dict, error_type):
    b, a = BASE_dict["LWUPT"], ALT_dict["LWUPT"]
    return a["avg"] - b["avg"], _quad(a[error_type], b[error_type])


# ---- shortwave (RadDecomp_functions.py:149-207) --------------------------------------------------------------------------
def calc_SW_DIRECT(BASE_dict, ALT_dict, error_type):
    """(SWUPTCLN - SWUPT)_BASE - (SWUPTCLN - SWUPT)_ALT"""
    cb, sb, ca, sa = BASE_dict["SWUPTCLN"], BASE_dict["SWUPT"], ALT_dict["SWUPTCLN"], ALT_dict["SWUPT"]
    eff = cb["avg"] - sb["avg"] - ca["avg"] + sa["avg"]
    return eff, _quad(cb[error_type], sb[error_type], ca[error_type], sa[error_type])


def calc_SW_INDIRECT(BASE_dict, ALT_dict, error_type):
    """SWUPTCLN_ALT_nA - SWUPTCLN_BASE_nA"""
    b, a = BASE_dict["SWUPTCLN_nA"], ALT_dict["SWUPTCLN_nA"]
    return a["avg"] - b["avg"], _quad(a[error_type], b[error_type])


def calc_SW_SEMIDIRECT(BASE_dict, ALT_dict, error_type):
    """SWUPTCLN_ALT - SWUPTCLN_BASE - SWUPT_ALT_nA + SWUPT_BASE_nA"""
    cb, sbn, ca, san = BASE_dict["SWUPTCLN"], BASE_dict["SWUPT_nA"], ALT_dict["SWUPTCLN"], ALT_dict["SWUPT_nA"]
    eff = ca["avg"] - cb["avg"] - san["avg"] + sbn["avg"]
    return eff, _quad(ca[error_type], cb[error_type], san[error_type], sbn[error_type])


# ---- longwave (RadDecomp_functions.py:216-267) ------------------------------------------------------------------------------
def calc_LW_INDIRECT(BASE_dict, ALT_dict, error_type, fixed=False):
    """(LWUPTC - LWUPT)_BASE_nA - (LWUPTC - LWUPT)_ALT_nA; by default with the reference's operands (see module docstring)."""
    cb, ca = BASE_dict["LWUPTC_nA"], ALT_dict["LWUPTC_nA"]
    lb = BASE_dict["LWUPT_nA"] if fixed else cb
    la = ALT_dict["LWUPT_nA"] if fixed else ca
    eff = cb["avg"] - lb["avg"] - ca["avg"] + la["avg"]
    return eff, _quad(cb[error_type], lb[error_type], ca[error_type], la[error_type])


def calc_LW_SEMIDIRECT(BASE_dict, ALT_dict, error_type):
    """(LWUPTC - LWUPT)_BASE - (LWUPTC - LWUPT)_ALT - [(LWUPTC - LWUPT)_BASE_nA - (LWUPTC - LWUPT)_ALT_nA]"""
    cb, lb, cbn, lbn = BASE_dict["LWUPTC"], BASE_dict["LWUPT"], BASE_dict["LWUPTC_nA"], BASE_dict["LWUPT_nA"]
    ca, la, can, lan = ALT_dict["LWUPTC"], ALT_dict["LWUPT"], ALT_dict["LWUPTC_nA"], ALT_dict["LWUPT_nA"]
    eff = cb["avg"] - lb["avg"] - ca["avg"] + la["avg"] - cbn["avg"] + lbn["avg"] + can["avg"] - lan["avg"]
    err = _quad(cb[error_type], lb[error_type], ca[error_type], la[error_type], cbn[error_type], lbn[error_type],
                can[error_type], lan[error_type])
    return eff, err


def decompose(BASE_dict, ALT_dict, error_type="standard_error", lw_indirect_fixed=False):
    """All seven terms at once: {name: (effect, error)}."""
    return {
        "Delta_S": calc_Delta_S(BASE_dict, ALT_dict, error_type),
        "Delta_L": calc_Delta_L(BASE_dict, ALT_dict, error_type),
        "SW_DIRECT": calc_SW_DIRECT(BASE_dict, ALT_dict, error_type),
        "SW_INDIRECT": calc_SW_INDIRECT(BASE_dict, ALT_dict, error_type),
        "SW_SEMIDIRECT": calc_SW_SEMIDIRECT(BASE_dict, ALT_dict, error_type),
        "LW_INDIRECT": calc_LW_INDIRECT(BASE_dict, ALT_dict, error_type, fixed=lw_indirect_fixed),
        "LW_SEMIDIRECT": calc_LW_SEMIDIRECT(BASE_dict, ALT_dict, error_type),
    }
