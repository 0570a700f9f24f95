"""The text files between the reference's two post-processing packages, written and read in the reference's own format.

The NCL package writes one file per variable and scenario, `<var>_domain_stats.txt` (`write_stats_data`,
analysis_scripts/NCL_extraction_package/data_extraction_library.ncl:428-574): a title line, the column header

    Time, Hour, min, x05, x25, median, x75, x95, max, avg, stddev, SE, morans_i, SE_corr, N

and one row per output time - the time label `(Mon-DD) HH` (create_local_time_strings, ncl:361-392), the hours since the
first time as `%6.2f` (calc_runtime_in_hours, ncl:398-420) and the 13 statistics of calc_standard_stats as `%7.4f`.  The
Python package reads `<DATADir>/<scen>/<var>_<dom>_stats.txt` with pandas (`load_Files`,
analysis_scripts/RadDecomp_analysis_package/RadDecomp_functions.py:95-112; `dom` = "domain" for the whole-domain files) and
uses the columns `avg`, `SE`, `SE_corr`.

Here the statistics come from the device (`Radiation.domain_statistics`: arc_rad_domain_stats / _percentiles / _morans_i), so
`write_stats_data` + the reference's own `load_Files` / `calc_*` - or the mirrors in this package - close the chain
radiation step -> domain statistics -> files -> decomposition without NCL.
"""
from __future__ import annotations

import os

import numpy as np

COL_HEAD = "Time, Hour, min, x05, x25, median, x75, x95, max, avg, stddev, SE, morans_i, SE_corr, N"
# column -> key of the statistics dict (decomposition.stats_from_sums / oracle ncl_stats.calc_standard_stats), in file order
# (stat_order = 2, 7, 5, 4, 6, 8, 3, 0, 1, 9, 10, 11, 12 of ncl:521)
FILE_ORDER = (("min", "min"), ("x05", "p05"), ("x25", "lower_quartile"), ("median", "median"), ("x75", "upper_quartile"), ("x95", "p95"),
              ("max", "max"), ("avg", "avg"), ("stddev", "stddev"), ("SE", "standard_error"), ("morans_i", "morans_i"),
              ("SE_corr", "corrected_standard_error"), ("N", "N"))
VAR_LIST = ("SWUPT", "LWUPT", "LWUPTC", "SWUPTCLN")       # load_Files, RadDecomp_functions.py:101
_MONTHS = ("Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec")


def create_local_time_strings(times, offset_hours=0.0):
    """`times`: sequence of datetime.datetime (the WRF output times); the NCL labels "(%c-%D) %H" = "(Jul-21) 06", shifted by
    use_local_time@offset hours when local time is asked for (ncl:361-392)."""
    import datetime
    out = []
    for t in times:
        t = t + datetime.timedelta(hours=float(offset_hours))
        out.append("(%s-%02d) %02d" % (_MONTHS[t.month - 1], t.day, t.hour))
    return out


def calc_runtime_in_hours(times):
    """Hours since the first output time (ncl:398-420)."""
    t0 = times[0]
    return np.array([(t - t0).total_seconds() / 3600.0 for t in times], dtype=np.float64)


def stats_file_name(var_name, region=None):
    """`<var>_domain_stats.txt`, or `<var>_<region>_domain_stats.txt` with region_select (ncl:510-514)."""
    return "%s_%sdomain_stats.txt" % (var_name, region + "_" if region else "")


def write_stats_data(output_directory, var_name, time_strings, hours, stats, units=None, region=None):
    """One single-level statistics file in the reference's format (the dimscount = 3 branch of write_stats_data).
    `stats`: one mapping per output time with the 13 statistics under the keys of decomposition.stats_from_sums
    (avg, stddev, min, max, median, lower_quartile, upper_quartile, p05, p95, standard_error, morans_i,
    corrected_standard_error, N).  Returns the path written."""
    if not (len(time_strings) == len(hours) == len(stats)):
        raise ValueError("write_stats_data: time_strings, hours and stats differ in length")
    lines = ["%s (%s)" % (var_name, units) if units else var_name, COL_HEAD]
    for ts, hr, st in zip(time_strings, hours, stats):
        row = ts + ", " + "%6.2f" % float(hr)
        for _, key in FILE_ORDER:
            row += ", " + "%7.4f" % float(st[key])
        lines.append(row)
    os.makedirs(output_directory, exist_ok=True)
    path = os.path.join(output_directory, stats_file_name(var_name, region))
    with open(path, "w") as fh:                 # asciiwrite: one string per line
        fh.write("\n".join(lines) + "\n")
    return path


def read_stats_file(path):
    """One statistics file -> {column: numpy array}, 'Time' as a list of labels, rows in file order."""
    with open(path) as fh:
        rows = [ln.rstrip("\n") for ln in fh if ln.strip()]
    cols = [c.strip() for c in rows[1].split(",")]
    data = [[c.strip() for c in r.split(",")] for r in rows[2:]]
    out = {"Time": [r[0] for r in data]}
    for q, c in enumerate(cols[1:], start=1):
        out[c] = np.array([float(r[q]) for r in data], dtype=np.float64)
    return out


def load_Files(DATADir, scen, alt_end, dom):
    """Mirror of RadDecomp_functions.load_Files (py:95-112): the four TOA variables of scenario `scen` and of its twin
    `scen + alt_end` (the run without aerosol-radiation interaction, '_nA'), each a pandas DataFrame indexed by 'Hour' - what
    the calc_* functions of the reference and of `decomposition` take (error_type 'SE' or 'SE_corr', the file's own columns)."""
    import pandas as pd
    out = {}
    for var in VAR_LIST:
        for suffix, directory in (("", scen), (alt_end, scen + alt_end)):
            cols = read_stats_file(os.path.join(DATADir, directory, "%s_%s_stats.txt" % (var, dom)))
            df = pd.DataFrame({k: v for k, v in cols.items() if k != "Hour"}, index=pd.Index(cols["Hour"], name="Hour"))
            out[var + suffix] = df
    return out


PLOT_VARIABLES = ("SWUPT", "SWUPTC", "SWUPTCLN", "SWDNT", "SWDNTC", "SWDNTCLN", "SWUPB", "SWUPBC", "SWUPBCLN", "SWDNB", "SWDNBC", "SWDNBCLN",
                  "LWUPT", "LWUPTC", "LWUPTCLN", "LWDNT", "LWDNTC", "LWDNTCLN", "LWUPB", "LWUPBC", "LWUPBCLN", "LWDNB", "LWDNBC", "LWDNBCLN")


def extract_name(var, clean_flag):
    """The variable actually read for `var`: a scenario without clean-sky output (clean_flag False: the runs without
    aerosol-radiation interaction) gets the all-aerosol variable in place of every *CLN one
    (load_2D_3D_variable_and_sample_at_given_altitudes, data_extraction_library.ncl:160-165)."""
    return var.replace("CLN", "") if (not clean_flag and "CLN" in var) else var


def extract_domain_averages(rad, dims, scenarios, output_root_directory, times, plot_variables=PLOT_VARIABLES, trim=5, regions=None,
                            local_time_offset=None, units="W m-2"):
    """The main loop of EXTRACT_domain_averages.ncl on in-memory fields instead of wrfout files: for every scenario and variable
    the 13 domain statistics of every output time (reduced on the device by `rad.domain_statistics`, 5-cell trim or region
    boxes as calculate_domain_stats cuts them) written to `<output_root>/<scenario>/<var>[_<region>]_domain_stats.txt`.

    scenarios: {name: (clean_flag, [ {wrf variable name: 2-D field} per output time ])} - host arrays or CUDA tensors;
    regions:   None, or {region name: (lon_start, lon_end, lat_start, lat_end)} index boxes (wrf_user_ll_to_ij results - 1).
    Returns {scenario: {(var, region or None): path}}."""
    labels = create_local_time_strings(times, local_time_offset or 0.0)
    hours = calc_runtime_in_hours(times)
    written = {}
    for scen, (clean_flag, per_time) in scenarios.items():
        if len(per_time) != len(times):
            raise ValueError("extract_domain_averages: scenario %s has %d times, expected %d" % (scen, len(per_time), len(times)))
        written[scen] = {}
        for region, box in (regions.items() if regions else [(None, None)]):
            rows = {v: [] for v in plot_variables}
            for fields in per_time:
                src = [extract_name(v, clean_flag) for v in plot_variables]
                missing = [n for n in src if n not in fields]
                if missing:
                    raise KeyError("extract_domain_averages: scenario %s lacks %s" % (scen, ", ".join(sorted(set(missing)))))
                uniq = sorted(set(src))
                st = rad.domain_statistics(dims, [fields[n] for n in uniq], names=uniq, trim=0 if box else trim, region=box)
                for v, n in zip(plot_variables, src):
                    rows[v].append(st[n])
            for v in plot_variables:
                written[scen][(v, region)] = write_stats_data(os.path.join(output_root_directory, scen), v, labels, hours, rows[v], units=units,
                                                              region=region)
    return written
