"""Host-side table preparation of the product (record parser + g-point reduction, csrc/tables.cpp) against the oracle's
independent restatement of sw_kgbNN / cmbgbNN (oracle/tables.cpp); malformed files are rejected with ARC_ERR_IO."""
import ctypes as C
import os

import numpy as np
import pytest

from wrfchem_arc_interactions_b200 import abi, radiation as R

SW = {16: ["absa", "absb", "selfref", "forref", "sfluxref"], 17: ["absa", "absb", "selfref", "forref", "sfluxref"],
      18: ["absa", "absb", "selfref", "forref", "sfluxref"], 20: ["absa", "absb", "absch4", "sfluxref"],
      23: ["absa", "raylg", "selfref"], 24: ["absa", "absb", "rayla", "raylb", "abso3a", "abso3b", "sfluxref"],
      25: ["absa", "abso3a", "abso3b", "raylg"], 26: ["raylg", "sfluxref"], 27: ["absa", "absb", "raylg"],
      28: ["absa", "absb", "sfluxref"], 29: ["absa", "absb", "absh2o", "absco2", "selfref", "forref"]}
LW = {1: ["absa", "absb", "ka_mn2", "kb_mn2", "fracrefa", "fracrefb", "selfref", "forref"], 3: ["absa", "absb", "ka_mn2o", "kb_mn2o", "fracrefa", "fracrefb"],
      5: ["absa", "absb", "ka_mo3", "ccl4"], 6: ["absa", "ka_mco2", "cfc11adj", "cfc12", "fracrefa"], 7: ["ka_mco2", "kb_mco2", "fracrefa", "fracrefb"],
      8: ["ka_mco2", "kb_mco2", "ka_mn2o", "kb_mn2o", "ka_mo3", "cfc12", "cfc22adj"], 9: ["ka_mn2o", "kb_mn2o"], 11: ["ka_mo2", "kb_mo2"],
      12: ["absa", "fracrefa"], 13: ["absa", "ka_mco2", "ka_mco", "kb_mo3", "fracrefb"], 15: ["absa", "ka_mn2"], 16: ["absa", "absb", "fracrefa", "fracrefb"]}


def host_table(lib, ktab, name, p_top=5000.0, kme=41):
    L = lib.lib
    L.arc_rad_host_table.restype = C.c_int
    L.arc_rad_host_table.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_float, C.c_float, C.c_int, C.c_char_p, abi.c_fp, C.c_int]
    args = (R.INLINE_TABLES.encode(), ktab[0].encode(), ktab[1].encode(), 1004.5, p_top, kme)
    n = L.arc_rad_host_table(*args, name.encode() if name else None, None, 0)
    if n <= 0 or name in (None, "lw_nlayers"):
        return n
    buf = np.zeros(n, np.float32)
    assert L.arc_rad_host_table(*args, name.encode(), buf.ctypes.data_as(abi.c_fp), n) == n
    return buf


def oracle_table(orc, kind, band, name):
    n = orc.lib.arc_oracle_table(kind, band, name.encode(), None, 0)
    assert n > 0, (kind, band, name)
    buf = np.zeros(n, np.float32)
    orc.lib.arc_oracle_table(kind, band, name.encode(), buf.ctypes.data_as(abi.c_fp), n)
    return buf


def test_reduced_tables_bit_exact(lib, orc, ktab):
    orc.init(5000.0, 41, ktab[0], ktab[1])
    checked = 0
    for band, names in SW.items():
        for nm in names:
            mine = host_table(lib, ktab, "sw%d.%s" % (band, "sflux" if nm == "sfluxref" else nm))
            ref = oracle_table(orc, 0, band - 15, nm)
            assert mine.shape == ref.shape and np.array_equal(mine, ref), ("sw", band, nm)
            checked += mine.size
    for band, names in LW.items():
        for nm in names:
            mine = host_table(lib, ktab, "lw%d.%s" % (band, nm))
            ref = oracle_table(orc, 1, band, nm)
            assert mine.shape == ref.shape and np.array_equal(mine, ref), ("lw", band, nm)
            checked += mine.size
    assert checked > 100000


def test_source_terms_sum_preserved(lib, ktab):
    """sfluxrefo / fracrefo are summed unweighted over merged g-points (SW:5103-5111): band totals are preserved."""
    s = host_table(lib, ktab, "sw16.sflux")
    assert abs(float(s.sum()) - 12.1096) < 1e-3
    f = host_table(lib, ktab, "lw1.fracrefa")
    assert abs(float(f.sum()) - 1.0) < 1e-5


@pytest.mark.parametrize("p_top,kme,expect", [(5000.0, 41, 53), (5000.0, 51, 63), (1000.0, 41, 43), (5000.0, 101, 113)])
def test_lw_nlayers_rule(lib, ktab, p_top, kme, expect):
    """LW:12861 nlayers = kme + nint(p_top*0.01/4) - 1, Fortran NINT rounds half away from zero (12.5 -> 13)."""
    L = lib.lib
    L.arc_rad_host_table.restype = C.c_int
    L.arc_rad_host_table.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_float, C.c_float, C.c_int, C.c_char_p, abi.c_fp, C.c_int]
    # a different key forces a rebuild: append a harmless "./" to the path
    sw = ktab[0].replace("RRTMG_SW_DATA", "./" * (kme % 7 + int(p_top) % 3) + "RRTMG_SW_DATA")
    n = L.arc_rad_host_table(R.INLINE_TABLES.encode(), sw.encode(), ktab[1].encode(), 1004.5, p_top, kme, b"lw_nlayers", None, 0)
    assert n == expect


def test_truncated_file_rejected(lib, ktab, tmp_path):
    L = lib.lib
    L.arc_rad_host_table.restype = C.c_int
    L.arc_rad_host_table.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_float, C.c_float, C.c_int, C.c_char_p, abi.c_fp, C.c_int]
    raw = open(ktab[0], "rb").read()
    bad = tmp_path / "RRTMG_SW_DATA"
    bad.write_bytes(raw[: len(raw) // 2])
    rc = L.arc_rad_host_table(R.INLINE_TABLES.encode(), str(bad).encode(), ktab[1].encode(), 1004.5, 5000.0, 41, None, None, 0)
    assert rc == -3           # ARC_ERR_IO
    assert b"RRTMG_SW_DATA band" in L.arc_rad_last_error()
    rc = L.arc_rad_host_table(R.INLINE_TABLES.encode(), str(tmp_path / "missing").encode(), ktab[1].encode(), 1004.5, 5000.0, 41, None, None, 0)
    assert rc == -3


def test_record_layout_sizes(ktab):
    """Record byte counts follow table T-K of SURVEY.md section 8a (e.g. band 16: 2 reals + 1 int + kao(9,5,13,16) + ...)."""
    import struct
    raw = open(ktab[0], "rb").read()
    n0 = struct.unpack("<i", raw[:4])[0]
    assert n0 == 4 * (2 + 1 + 9 * 5 * 13 * 16 + 5 * 47 * 16 + 10 * 16 + 3 * 16 + 16)
    assert struct.unpack("<i", raw[4 + n0:8 + n0])[0] == n0


def _to_big_endian(src, dst):
    """Rewrite a little-endian Fortran sequential-unformatted file of 4-byte items as WRF's big-endian flavour
    (-fconvert=big-endian: record markers and REAL*4 / INTEGER*4 payload byte-swapped)."""
    raw = np.fromfile(src, dtype="<u4")
    raw.astype(">u4").tofile(dst)


def test_big_endian_files_give_the_same_tables(lib, ktab, tmp_path):
    """The data files WRF ships are big-endian (WRF is compiled with -fconvert=big-endian); the reader detects the byte
    order from the first record's markers (csrc/tables.cpp FileCursor)."""
    be = (str(tmp_path / "RRTMG_SW_DATA"), str(tmp_path / "RRTMG_LW_DATA"))
    _to_big_endian(ktab[0], be[0]); _to_big_endian(ktab[1], be[1])
    assert open(be[0], "rb").read(4) != open(ktab[0], "rb").read(4)
    assert host_table(lib, be, None) == 0
    n = 0
    for name in ("sw16.absa", "sw17.absb", "sw24.rayla", "sw29.absco2", "sw22.sflux", "lw1.ka_mn2", "lw3.absa", "lw5.absb", "lw13.ka_mco", "lw8.cfc22adj",
                 "lw16.fracrefa"):
        a, b = host_table(lib, ktab, name), host_table(lib, be, name)
        assert a.shape == b.shape and np.array_equal(a, b), name
        n += a.size
    assert n > 40000
    # mixed byte order inside one file is an error, not garbage
    bad = str(tmp_path / "mixed")
    raw = bytearray(open(be[1], "rb").read())
    raw[0:4] = raw[0:4][::-1]
    open(bad, "wb").write(raw)
    assert host_table(lib, (ktab[0], bad), None) == -3          # -ARC_ERR_IO
