"""Full-size comparisons of the CUDA path with the oracle (all host threads), shared by tests/test_gpu_parity.py and
tools/parity_report.py (which writes profiles/r2_parity.md).  One tolerance for every column: 1e-4 relative or 0.01 W/m2
absolute for every flux output, 1e-4 or 0.01 K/day for the heating rates."""
import numpy as np

from wrfchem_arc_interactions_b200 import synth

# name, ni, nj, nk, synth kwargs, what it stands for
FULL_CASES = [
    ("C1", 32, 32, 40, {}, "BASELINE config 1, complete: 1,024 columns x 40 levels"),
    ("C2", 425, 300, 50, {}, "BASELINE config 2, complete: 425x300 = 127,500 columns x 50 levels, 40% cloudy, 25% night"),
    ("C3 slice", 400, 60, 60, {}, "config 3's grid (60 levels; SW 61 / LW 73 layers), 24,000 of its 4 M columns"),
    ("C4 slice", 425, 50, 50, dict(cloudy_frac=1.0, with_re=True), "config 4: every column cloudy, re_cloud / re_ice / re_snow given (inflg 5), 21,250 columns"),
    ("C5 slice", 400, 52, 100, {}, "config 5's grid (100 levels; SW 101 / LW 113 layers), 20,800 of its 1 M columns"),
]


def within(a, b, rel=1e-4, ab=0.01):
    err = np.abs(a - b)
    return (err <= rel * np.abs(b)) | (err <= ab)


def per_column(dom, a):
    h = dom["halo"]
    if h:
        a = a[h:-h, ..., h:-h] if a.ndim == 3 else a[h:-h, h:-h]
    return a.transpose(0, 2, 1).reshape(-1, a.shape[1]) if a.ndim == 3 else a.reshape(-1)


def compare_outputs(dom, og, oo):
    """Every output of one adapter against the oracle: (columns compared, columns out of tolerance, worst absolute deviation
    in W/m2, worst relative deviation among cells that deviate by more than 0.01 W/m2, worst heating-rate deviation, K/day)."""
    nz = dom["nk"]
    bad = np.zeros(dom["ni"] * dom["nj"], bool)
    worst_abs = worst_rel = worst_hr = 0.0
    for k in og:
        a, b = per_column(dom, og[k]).astype(np.float64), per_column(dom, oo[k]).astype(np.float64)
        if k.startswith("rthraten"):
            a, b = a[:, :nz] * 86400.0, b[:, :nz] * 86400.0          # K/s -> K/day (the Exner factor is in both)
            worst_hr = max(worst_hr, float(np.abs(a - b).max()))
        elif k == "coszr":
            assert np.array_equal(a, b)
            continue
        else:
            err = np.abs(a - b)
            worst_abs = max(worst_abs, float(err.max()))
            big = err > 0.01
            if big.any():
                worst_rel = max(worst_rel, float((err[big] / np.maximum(np.abs(b[big]), 1e-30)).max()))
        ok = within(a, b)
        bad |= ~(ok if ok.ndim == 1 else ok.all(axis=1))
    nbits = 0
    for k in og:
        a, b = per_column(dom, og[k]), per_column(dom, oo[k])
        if a.ndim == 2:
            a, b = a[:, :nz + (0 if k.startswith("rthraten") else 2)], b[:, :nz + (0 if k.startswith("rthraten") else 2)]
        nbits += int((np.ascontiguousarray(a).view(np.uint32) != np.ascontiguousarray(b).view(np.uint32)).sum())
    return dict(columns=int(bad.size), out_of_tolerance=int(bad.sum()), worst_abs=worst_abs, worst_rel=worst_rel, worst_hr=worst_hr,
                cells_not_bit_exact=nbits)


def run_case(case, lib, orc_mt, ktab, run_pair, init):
    name, ni, nj, nk, kw, _ = case
    dom = synth.make_domain(ni, nj, nk, seed=synth.SEED + 77, **kw)
    init(lib, dom, ktab); init(orc_mt, dom, ktab)
    res = {}
    for which in ("sw", "lw"):
        og = run_pair(which, lib, dom)
        oo = run_pair(which, orc_mt, dom)
        res[which] = compare_outputs(dom, og, oo)
        if which == "sw":
            res[which]["sunlit"] = int((dom["xcoszen"] > 0).sum())
    return res
