"""radconst / calc_coszen of radiation_driver (module_radiation_driver.F:2595-2666): oracle known answers on the CPU."""
import ctypes as C

import numpy as np

from wrfchem_arc_interactions_b200 import abi

DEGRAD, DPD = np.float32(3.1415926 / 180.0), np.float32(360.0 / 365.0)


def radconst(L, prefix, julian):
    fn = getattr(L, prefix + "radconst")
    fn.restype = None
    fn.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, abi.c_fp, abi.c_fp]
    d, s = np.zeros(1, np.float32), np.zeros(1, np.float32)
    fn(0.0, float(julian), float(DEGRAD), float(DPD), abi.fptr(d), abi.fptr(s))
    return float(d[0]), float(s[0])


def test_radconst_known_answers(orc, lib):
    for L, p in ((orc.lib, "arc_oracle_"), (lib.lib, "arc_rad_")):
        d_eq, _ = radconst(L, p, 80.0)                       # vernal equinox: declination 0
        assert abs(d_eq) < 1e-6
        d_sol, _ = radconst(L, p, 80.0 + 365.0 / 4.0)        # ~June solstice: +23.5 degrees
        assert abs(np.degrees(d_sol) - 23.5) < 0.05
        d_win, _ = radconst(L, p, 355.0)
        assert np.degrees(d_win) < -23.0
        _, s_jan = radconst(L, p, 3.0); _, s_jul = radconst(L, p, 185.0)
        assert 1410 < s_jan < 1420 and 1320 < s_jul < 1330    # perihelion / aphelion: 1370 * (1 +- 0.034)
    # the product's host arithmetic is the oracle's, bit for bit
    for jd in (1.0, 79.5, 80.0, 200.25, 365.0):
        assert radconst(orc.lib, "arc_oracle_", jd) == radconst(lib.lib, "arc_rad_", jd)


def test_calc_coszen_oracle(orc):
    L = orc.lib
    L.arc_oracle_calc_coszen.restype = C.c_int
    L.arc_oracle_calc_coszen.argtypes = [C.POINTER(abi.ArcDims)] + [C.c_float] * 5 + [abi.c_fp] * 4
    ni, nj = 36, 9
    dims = abi.make_dims(dict(ids=1, ide=ni + 1, jds=1, jde=nj + 1, kds=1, kde=2, ims=1, ime=ni, jms=1, jme=nj, kms=1, kme=2,
                              its=1, ite=ni, jts=1, jte=nj, kts=1, kte=1))
    lon = np.tile(np.linspace(-180, 170, ni, dtype=np.float32), (nj, 1)); lat = np.tile(np.linspace(-80, 80, nj, dtype=np.float32)[:, None], (1, ni))
    lon, lat = np.ascontiguousarray(lon), np.ascontiguousarray(lat)
    cz, hr = np.zeros_like(lon), np.zeros_like(lon)
    declin, _ = radconst(L, "arc_oracle_", 80.0)
    L.arc_oracle_calc_coszen(C.byref(dims), 80.0, 720.0, 0.0, declin, float(DEGRAD), abi.fptr(lon), abi.fptr(lat), abi.fptr(cz), abi.fptr(hr))
    # 12 UTC at the equinox: the sun is overhead near lon 0 (equation of time ~ -7.5 min -> ~2 deg), cos(zenith) = cos(lat) there
    j0 = nj // 2
    i_noon = int(np.argmax(cz[j0]))
    assert abs(lon[j0, i_noon]) <= 10.0 and cz[j0].max() > 0.99
    assert np.allclose(cz[:, i_noon], np.cos(np.radians(lat[:, i_noon])), atol=2e-3)
    assert (cz < 0).mean() > 0.35 and (cz > 0).mean() > 0.35          # half the globe is dark
