#!/usr/bin/env python
"""Golden fixture for the statistics text files: four scenario directories written by OUR write_stats_data, read back by the
REFERENCE's own load_Files (/root/reference/analysis_scripts/RadDecomp_analysis_package/RadDecomp_functions.py:95-112, imported
in the development container) and pushed through its calc_* functions with error_type = 'SE_corr', the setting of
RadDecomp_DiurnalAvg_timeplot.py:102.  The reference cannot travel to the GPU box: what it parsed and computed is committed as
tests/golden/statsfile_golden.npz together with this script.  (matplotlib is not installed: stubbed, as in
make_golden_raddecomp.py.)

    python tests/golden/make_golden_statsfile.py
"""
import datetime
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/analysis_scripts/RadDecomp_analysis_package"
SCENARIOS = ("BASE", "BASE_nA", "ALT", "ALT_nA")
NT = 9


def make_stats(seed=20160650):
    """{scenario: {var: [13-statistics dict per time]}} - synthetic, but with the structure of real output."""
    from wrfchem_arc_interactions_b200 import stats_files as SF
    rng = np.random.default_rng(seed)
    out = {}
    for s, scen in enumerate(SCENARIOS):
        out[scen] = {}
        for var in SF.VAR_LIST:
            base = {"SWUPT": 212.0, "LWUPT": 263.0, "LWUPTC": 280.0, "SWUPTCLN": 204.0}[var] + 1.7 * s
            rows = []
            for t in range(NT):
                x = np.sort(base + rng.normal(0.0, 25.0, 400))
                sd = float(x.std(ddof=1)); mi = float(rng.uniform(0.3, 0.95)); se = sd / 20.0
                rows.append({"avg": float(x.mean()), "stddev": sd, "min": float(x[0]), "max": float(x[-1]), "median": float(x[200]),
                             "lower_quartile": float(x[100]), "upper_quartile": float(x[299]), "p05": float(x[20]), "p95": float(x[379]),
                             "standard_error": se, "morans_i": mi, "corrected_standard_error": se * mi, "N": 400.0})
            out[scen][var] = rows
    return out


def write_all(directory, stats):
    from wrfchem_arc_interactions_b200 import stats_files as SF
    times = [datetime.datetime(2012, 7, 21, 0) + datetime.timedelta(hours=3 * t) for t in range(NT)]
    for scen in SCENARIOS:
        for var in SF.VAR_LIST:
            SF.write_stats_data(os.path.join(directory, scen), var, SF.create_local_time_strings(times), SF.calc_runtime_in_hours(times),
                                stats[scen][var], units="W m-2")


def main():
    mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot"); mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = plt
    sys.path.insert(0, REF)
    import RadDecomp_functions as RD
    d = tempfile.mkdtemp()
    write_all(d, make_stats())
    B, A = RD.load_Files(d, "BASE", "_nA", "domain"), RD.load_Files(d, "ALT", "_nA", "domain")
    out = {}
    for name, dd in (("BASE", B), ("ALT", A)):
        for k, df in dd.items():
            out["parsed/%s/%s/index" % (name, k)] = df.index.to_numpy(dtype=np.float64)
            for col in ("avg", "SE", "SE_corr", "median", "N"):
                out["parsed/%s/%s/%s" % (name, k, col)] = df[col].to_numpy(dtype=np.float64)
    for fn in ("calc_Delta_S", "calc_Delta_L", "calc_SW_DIRECT", "calc_SW_INDIRECT", "calc_SW_SEMIDIRECT", "calc_LW_INDIRECT", "calc_LW_SEMIDIRECT"):
        eff, err = getattr(RD, fn)(B, A, "SE_corr")
        out["calc/%s/effect" % fn] = np.asarray(eff, dtype=np.float64); out["calc/%s/error" % fn] = np.asarray(err, dtype=np.float64)
    out["text/BASE/SWUPT"] = np.frombuffer(open(os.path.join(d, "BASE", "SWUPT_domain_stats.txt"), "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "statsfile_golden.npz"), **out)
    print("wrote statsfile_golden.npz:", len(out), "arrays;", open(os.path.join(d, "BASE", "SWUPT_domain_stats.txt")).read().splitlines()[:3])


if __name__ == "__main__":
    main()
