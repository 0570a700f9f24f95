"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's NCL domain statistics, line by line.
PARITY UNPINNED: NCL is not available in this image and the reference ships no sample output; the restatement follows
analysis_scripts/NCL_extraction_package/misc_stats_library.ncl and is pinned only by the structural checks in
tests/test_ncl_stats_cpu.py.  Only tests/ may import this module.

  calc_morans_i_2D      misc_stats_library.ncl:196-371   (options as calc_standard_stats passes them: neighbour, manhattan)
  calc_boxplot_stats    misc_stats_library.ncl:145-189   (5th, 25th, 50th, 75th, 95th percentile = sorted(x)[round(.01 p (N-1))])
  calc_standard_stats   misc_stats_library.ncl:396-461   (all 13 statistics)
  calculate_domain_stats' domain trim  data_extraction_library.ncl:318-322   (trim cells cut from every edge)
"""
import numpy as np


def calc_morans_i_2D(data, neighbour=True, manhattan=True):
    """Literal restatement for neighbour mode: eight raveled-index displacements with wrap-around (ncl:273-281), Manhattan
    or Euclidean distance of the two cells (ncl:320-329), weight 1 where 0 < dist < 2 (ncl:333-335), autocorrelation and
    weight sums in double (ncl:350-357), Morans_I = tofloat(Auto_sum / (W_sum * stddev^2)) (ncl:366)."""
    assert neighbour, "only the mode calc_standard_stats uses is restated"
    data = np.asarray(data, dtype=np.float32)
    X_mean = np.float32(data.astype(np.float64).mean())
    X_sttdev = np.float32(data.astype(np.float64).std(ddof=1))
    X_diff = (data - X_mean).astype(np.float32)
    if np.all(X_diff == 0.0):
        return np.float32(0.0)
    n_j, n_i = data.shape
    i_ind = np.broadcast_to(np.arange(n_i, dtype=np.float32)[None, :], (n_j, n_i)).ravel()
    j_ind = np.broadcast_to(np.arange(n_j, dtype=np.float32)[:, None], (n_j, n_i)).ravel()
    X_diff_1D = X_diff.ravel()
    N_tot = X_diff_1D.size
    N_span = np.arange(N_tot)
    m_arr = [1, n_i - 1, n_i, n_i + 1, N_tot - n_i - 1, N_tot - n_i, N_tot - n_i + 1, N_tot - 1]
    W_sum = 0.0
    Auto_sum = 0.0
    for m in m_arr:
        if m == 0:
            continue
        N_mod = np.mod(N_span + m, N_tot)
        if manhattan:
            dist = np.sqrt((i_ind[N_span] - i_ind[N_mod]) ** 2) + np.sqrt((j_ind[N_span] - j_ind[N_mod]) ** 2)
        else:
            dist = np.sqrt((i_ind[N_span] - i_ind[N_mod]) ** 2 + (j_ind[N_span] - j_ind[N_mod]) ** 2)
        w = np.where((dist < 2) & (dist > 0), 1.0, 0.0)
        if w.sum() < 1e-2:
            continue
        Auto_sum += float(np.sum(w * X_diff_1D[N_span].astype(np.float64) * X_diff_1D[N_mod].astype(np.float64)))
        W_sum += float(w.sum())
    return np.float32(Auto_sum / (W_sum * float(X_sttdev * X_sttdev)))


def ncl_round(x):
    """NCL round(x, 3): nearest integer, halves away from zero."""
    x = float(x)
    return int(np.floor(x + 0.5)) if x >= 0 else -int(np.floor(-x + 0.5))


def calc_boxplot_stats(data):
    """ncl:145-189: qsort, then for perc_point = 5, 25, 50, 75, 95: pt_x = round(.01*perc_point*(numel-1), 3) (single
    precision, left to right), negative pt_x -> 0, stats(j) = oneD_data(pt_x)."""
    x = np.sort(np.asarray(data, dtype=np.float32).ravel())
    numel = x.size
    out = []
    for p in (5.0, 25.0, 50.0, 75.0, 95.0):
        pt_x = ncl_round(np.float32(np.float32(np.float32(0.01) * np.float32(p)) * np.float32(numel - 1)))
        out.append(x[max(pt_x, 0)])
    return np.array(out, np.float32)


def calc_standard_stats(data, trim=0):
    """All 13 statistics of ncl:432-458 in the reference's order of meaning: avg, stddev (N-1), min, max, median, lower / upper
    quartile, 5th / 95th percentile, standard error, Moran's I, corrected standard error = SE * I, N.  `trim` cells are cut
    from every edge first, as calculate_domain_stats does with domain_trim@trim (data_extraction_library.ncl:318-322)."""
    data = np.asarray(data, dtype=np.float32)
    if trim:
        data = data[trim:data.shape[0] - trim, trim:data.shape[1] - trim]
    x = data.astype(np.float64).ravel()
    sd = x.std(ddof=1)
    se = sd / np.sqrt(x.size)
    mi = float(calc_morans_i_2D(data))
    bp = calc_boxplot_stats(data)
    return {"avg": x.mean(), "stddev": sd, "min": float(x.min()), "max": float(x.max()), "median": float(bp[2]), "lower_quartile": float(bp[1]),
            "upper_quartile": float(bp[3]), "p05": float(bp[0]), "p95": float(bp[4]), "standard_error": se, "morans_i": mi,
            "corrected_standard_error": se * mi, "N": x.size}
