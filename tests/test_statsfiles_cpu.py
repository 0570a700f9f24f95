"""The statistics text files between the reference's NCL and Python packages (the data format either side of
the decomposition): our writer produces the format write_stats_data defines (data_extraction_library.ncl:428-574); what the
REFERENCE's own load_Files and calc_* made of those files is a committed golden (tests/golden/make_golden_statsfile.py)."""
import datetime
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from wrfchem_arc_interactions_b200 import decomposition as D, stats_files as SF

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden_statsfile as G  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "statsfile_golden.npz"))
FNS = ("calc_Delta_S", "calc_Delta_L", "calc_SW_DIRECT", "calc_SW_INDIRECT", "calc_SW_SEMIDIRECT", "calc_LW_INDIRECT", "calc_LW_SEMIDIRECT")


@pytest.fixture(scope="module")
def written(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("stats"))
    G.write_all(d, G.make_stats())
    return d


def test_file_format_is_the_ncl_format(written):
    """Title line, the NCL column header, '(Mon-DD) HH' labels, %6.2f hours, %7.4f statistics in stat_order - byte for byte the
    file the golden was made from."""
    text = open(os.path.join(written, "BASE", "SWUPT_domain_stats.txt"), "rb").read()
    assert text == GOLD["text/BASE/SWUPT"].tobytes()
    lines = text.decode().splitlines()
    assert lines[0] == "SWUPT (W m-2)" and lines[1] == SF.COL_HEAD and len(lines) == 2 + G.NT
    assert lines[2].startswith("(Jul-21) 00,   0.00, ") and lines[3].startswith("(Jul-21) 03,   3.00, ")
    assert all(len(ln.split(",")) == 15 for ln in lines[1:])
    assert SF.stats_file_name("LWUPT", "NChina") == "LWUPT_NChina_domain_stats.txt"
    t = [datetime.datetime(2012, 12, 31, 22), datetime.datetime(2013, 1, 1, 1)]
    assert SF.create_local_time_strings(t, offset_hours=8) == ["(Jan-01) 06", "(Jan-01) 09"] and list(SF.calc_runtime_in_hours(t)) == [0.0, 3.0]


def test_load_files_equals_the_reference_parser(written):
    B, A = SF.load_Files(written, "BASE", "_nA", "domain"), SF.load_Files(written, "ALT", "_nA", "domain")
    for name, dd in (("BASE", B), ("ALT", A)):
        assert sorted(dd) == sorted(v + s for v in SF.VAR_LIST for s in ("", "_nA"))
        for k, df in dd.items():
            assert np.array_equal(df.index.to_numpy(dtype=np.float64), GOLD["parsed/%s/%s/index" % (name, k)])
            for col in ("avg", "SE", "SE_corr", "median", "N"):
                assert np.array_equal(df[col].to_numpy(dtype=np.float64), GOLD["parsed/%s/%s/%s" % (name, k, col)]), (name, k, col)


def test_decomposition_from_files_equals_the_reference(written):
    """files -> load_Files -> calc_* with error_type 'SE_corr' (RadDecomp_DiurnalAvg_timeplot.py:102): effects and errors equal
    what the reference's functions returned on the same files, bit for bit."""
    B, A = SF.load_Files(written, "BASE", "_nA", "domain"), SF.load_Files(written, "ALT", "_nA", "domain")
    for fn in FNS:
        eff, err = getattr(D, fn)(B, A, "SE_corr")
        assert np.array_equal(np.asarray(eff, dtype=np.float64), GOLD["calc/%s/effect" % fn]), fn
        assert np.array_equal(np.asarray(err, dtype=np.float64), GOLD["calc/%s/error" % fn]), fn


def test_device_statistics_dicts_feed_the_writer(tmp_path):
    """The dicts stats_from_sums builds from arc_rad_domain_stats / _percentiles / _morans_i outputs carry every column the file needs."""
    sums = np.array([[2.0e5, 4.3e7, 1000.0, 120.0, 310.0]])
    st = D.stats_from_sums(sums, names=["SWUPT"], morans_i=[0.8], percentiles=[[200.0, 180.0, 220.0, 150.0, 260.0]])["SWUPT"]
    p = SF.write_stats_data(str(tmp_path), "SWUPT", ["(Jul-21) 00"], [0.0], [st])
    row = SF.read_stats_file(p)
    assert row["avg"][0] == 200.0 and row["N"][0] == 1000.0 and row["median"][0] == 200.0 and row["x95"][0] == 260.0
    assert abs(row["SE_corr"][0] - 0.8 * row["SE"][0]) < 1e-4
    with pytest.raises(ValueError):
        SF.write_stats_data(str(tmp_path), "SWUPT", ["a", "b"], [0.0], [st])


class _HostStats:
    """Stand-in for Radiation.domain_statistics on the CPU: the restated NCL statistics of oracle/ncl_stats.py."""
    def domain_statistics(self, dims, fields, names=None, morans=True, percentiles=True, trim=0, region=None):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ncl_stats as N
        out = {}
        for n, f in zip(names, fields):
            x = f if region is None else f[region[2]:region[3] + 1, region[0]:region[1] + 1]
            out[n] = N.calc_standard_stats(x, trim=0 if region is not None else trim)
        return out


def test_extract_domain_averages_driver(tmp_path):
    """EXTRACT_domain_averages.ncl's loop: every variable of every scenario to its file; a scenario without clean-sky output gets
    the all-aerosol field for *CLN (ncl:160-165); 5-cell trim; region boxes name their files <var>_<region>_domain_stats.txt."""
    rng = np.random.default_rng(3)
    times = [datetime.datetime(2012, 7, 21, 0), datetime.datetime(2012, 7, 21, 3)]
    mk = lambda names: [{n: rng.normal(200.0, 20.0, (24, 30)).astype(np.float32) for n in names} for _ in times]
    full = mk(SF.PLOT_VARIABLES)
    nocln = mk([v for v in SF.PLOT_VARIABLES if "CLN" not in v])
    out = SF.extract_domain_averages(_HostStats(), None, {"Basecase": (True, full), "Basecase_nA": (False, nocln)}, str(tmp_path), times)
    assert len(out["Basecase"]) == 24 and os.path.basename(out["Basecase"][("SWUPTCLN", None)]) == "SWUPTCLN_domain_stats.txt"
    a = SF.read_stats_file(out["Basecase_nA"][("SWUPTCLN", None)]); b = SF.read_stats_file(out["Basecase_nA"][("SWUPT", None)])
    assert np.array_equal(a["avg"], b["avg"]) and a["Time"] == ["(Jul-21) 00", "(Jul-21) 03"] and list(a["Hour"]) == [0.0, 3.0]
    assert a["N"][0] == (24 - 10) * (30 - 10)                                   # domain_trim@trim = 5
    want = float(np.float64(full[1]["LWDNB"][5:-5, 5:-5].astype(np.float64).mean()))
    assert abs(SF.read_stats_file(out["Basecase"][("LWDNB", None)])["avg"][1] - want) < 6e-5
    reg = SF.extract_domain_averages(_HostStats(), None, {"Basecase": (True, full)}, str(tmp_path / "r"), times, plot_variables=("SWUPT",),
                                     regions={"ENG": (4, 11, 2, 9)})
    r = SF.read_stats_file(reg["Basecase"][("SWUPT", "ENG")])
    assert os.path.basename(reg["Basecase"][("SWUPT", "ENG")]) == "SWUPT_ENG_domain_stats.txt" and r["N"][0] == 64
    with pytest.raises(KeyError):
        SF.extract_domain_averages(_HostStats(), None, {"x": (True, nocln)}, str(tmp_path / "e"), times)
