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


def test_cal_cldfra1_oracle_known_answers(orc):
    """cal_cldfra1 (module_radiation_driver.F:2886-3122) restated in the oracle: no condensate -> 0 (flag 1); condensate at or above
    saturation -> 1 (flag 2); subsaturated -> (RH)^0.25 (1 - exp(-100 q / subsat^0.49)) with the 0.01 floor (flag 3); the
    ice / water weighting of the saturation mixing ratio; OPTIONAL flags not PRESENT -> 0."""
    from wrfchem_arc_interactions_b200 import synth
    dom = synth.make_domain(12, 5, 40, seed=3, cloudy_frac=1.0)
    shp = dom["t3d"].shape
    cf = np.full(shp, -9.0, np.float32); fl = np.full(shp, -9, np.int32)
    p = dom["p3d"]; t = dom["t3d"]
    orc.cal_cldfra1(dom["dims"], cf, dom["qv3d"], dom["qc3d"], dom["qi3d"], dom["qs3d"], t, p, cldfra1_flag=fl)
    tile = (slice(None), slice(0, 40), slice(None))
    assert np.all(cf[:, 40, :] == -9.0)                              # the level kme is not part of the tile
    c, f = cf[tile], fl[tile]
    qcld = (dom["qi3d"] + dom["qc3d"] + dom["qs3d"])[tile]
    assert np.all(c[qcld < 1e-12] == 0) and np.all(f[qcld < 1e-12] == 1)
    assert set(np.unique(f)) <= {1, 2, 3} and (f == 3).any()
    assert np.all((c == 0) | (c >= 0.01)) and c.max() <= 1.0
    assert np.all(c[f == 2] == 1.0)
    # closed form of one partially cloudy cell, in float64 (the oracle runs float32)
    jj, kk, ii = [x[0] for x in np.nonzero((f == 3) & (c > 0.05) & (c < 0.95))]
    tk, pp = float(t[jj, kk, ii]), float(p[jj, kk, ii])
    esw = 610.78 * np.exp(17.2693882 * (tk - 273.15) / (tk - 35.86)); esi = 610.78 * np.exp(21.8745584 * (tk - 273.15) / (tk - 7.66))
    qvsw, qvsi = 287.0 / 461.6 * esw / (pp - esw), 287.0 / 461.6 * esi / (pp - esi)
    qi, qc, qs, qv = (float(dom[k][jj, kk, ii]) for k in ("qi3d", "qc3d", "qs3d", "qv3d"))
    w = (qi + qs) / (qi + qc + qs)
    qvs = (1 - w) * qvsw + w * qvsi
    want = (qv / qvs) ** 0.25 * (1 - np.exp(max(-6.9, -100.0 * (qi + qc + qs) / max(1e-10, qvs - qv) ** 0.49)))
    assert abs(float(c[jj, kk, ii]) - want) < 2e-5
    # OPTIONAL flags not PRESENT: zero everywhere
    cf2 = np.full(shp, -9.0, np.float32)
    orc.cal_cldfra1(dom["dims"], cf2, dom["qv3d"], dom["qc3d"], dom["qi3d"], dom["qs3d"], t, p, F_QI=None)
    assert np.all(cf2[tile] == 0)
    # qc only (MP options 1 / 3): ice saturation at and below freezing
    cf3 = np.zeros(shp, np.float32)
    orc.cal_cldfra1(dom["dims"], cf3, dom["qv3d"], dom["qc3d"], None, None, t, p, F_QI=False, F_QS=False)
    assert (cf3[tile] > 0).any()


def test_cal_cldfra2_oracle(orc):
    """cal_cldfra2 (module_radiation_driver.F:2801-2874): 1 where QC + QI > 1e-6; QC alone when F_QI is false; 0 when F_QC is false."""
    from wrfchem_arc_interactions_b200 import synth
    dom = synth.make_domain(12, 5, 40, seed=4, cloudy_frac=1.0)
    qc, qi = dom["qc3d"], dom["qi3d"]
    tile = (slice(None), slice(0, 40), slice(None))
    cf = np.full(qc.shape, -9.0, np.float32)
    orc.cal_cldfra2(dom["dims"], cf, qc, qi)
    assert np.all(cf[:, 40, :] == -9.0)
    assert np.array_equal(cf[tile], ((qc + qi)[tile] > np.float32(1e-6)).astype(np.float32)) and 0 < cf[tile].mean() < 1
    orc.cal_cldfra2(dom["dims"], cf, qc, qi, F_QI=False)
    assert np.array_equal(cf[tile], (qc[tile] > np.float32(1e-6)).astype(np.float32))
    orc.cal_cldfra2(dom["dims"], cf, qc, qi, F_QC=False)
    assert np.all(cf[tile] == 0)


def ozone_case(ni=14, nj=6, nk=40, levsiz=59, seed=11):
    """Synthetic CAM-style ozone climatology: 12 months on `levsiz` pressure levels (top down), and a model pressure field that
    reaches above the data top and below the data bottom in some columns."""
    from wrfchem_arc_interactions_b200 import synth
    rng = np.random.default_rng(seed)
    dom = synth.make_domain(ni, nj, nk, seed=seed)
    pin = np.geomspace(30.0, 99000.0, levsiz).astype(np.float32)                  # Pa, top down
    prof = (8e-6 * np.exp(-((np.log(pin) - np.log(1000.0)) / 1.2) ** 2) + 3e-8).astype(np.float32)
    ozmixm = (prof[None, None, :, None] * (1 + 0.2 * rng.random((12, nj, levsiz, ni)))).astype(np.float32)
    p = dom["p3d"].copy()
    p[:, :, 0] *= np.float32(1.08)                                                 # one column row-edge below the data bottom
    p[0, nk - 1, :] = 20.0                                                         # above the data top
    p[1, nk - 1, 3] = pin[0]                                                       # exactly on the top data level
    return dom, pin, ozmixm, np.ascontiguousarray(p)


def test_ozn_time_int_oracle(orc):
    """ozn_time_int (DRV:3993-4098): on a mid-month day the field is that month's; between two mid-month days it is the linear
    blend; the December-January wrap uses both ends of the year."""
    dom, pin, ozmixm, p = ozone_case()
    nj, levsiz, ni = ozmixm.shape[1:]
    ozt = np.zeros((nj, levsiz, ni), np.float32)
    orc.ozn_time_int(dom["dims"], 0, 44.0, ozmixm, ozt, levsiz, 12)               # JULIAN + 1 = 45 = date_oz(2): all of February
    assert np.array_equal(ozt, ozmixm[1])
    orc.ozn_time_int(dom["dims"], 0, 59.0, ozmixm, ozt, levsiz, 12)               # day 60: half way between 45 and 75
    assert np.allclose(ozt, 0.5 * (ozmixm[1] + ozmixm[2]), rtol=1e-6)
    orc.ozn_time_int(dom["dims"], 0, 359.5, ozmixm, ozt, levsiz, 12)              # day 360.5: December -> January wrap
    f2 = (360.5 - 350.0) / 31.0
    assert np.allclose(ozt, (1 - f2) * ozmixm[11] + f2 * ozmixm[0], rtol=2e-6)
    orc.ozn_time_int(dom["dims"], 0, 4.25, ozmixm, ozt, levsiz, 12)               # day 5.25: January side of the wrap
    f2 = (5.25 + 365.0 - 350.0) / 31.0
    assert np.allclose(ozt, (1 - f2) * ozmixm[11] + f2 * ozmixm[0], rtol=2e-6)


def test_ozn_p_int_oracle(orc):
    """ozn_p_int (DRV:4100-4234) against numpy's linear interpolation in pressure inside the data range, the p / pin(1) scaling
    above the top data level and the held value below the bottom one."""
    dom, pin, ozmixm, p = ozone_case()
    nj, levsiz, ni = ozmixm.shape[1:]
    nk = dom["nk"]
    ozt = np.ascontiguousarray(ozmixm[4])
    o3 = np.full(p.shape, -1.0, np.float32)
    orc.ozn_p_int(dom["dims"], p, pin, levsiz, ozt, o3)
    assert np.all(o3[:, nk, :] == -1.0) and np.all(o3[:, :nk, :] > 0)
    for j in range(nj):
        for i in range(ni):
            pm = p[j, :nk, i].astype(np.float64)
            want = np.interp(pm, pin.astype(np.float64), ozt[j, :, i].astype(np.float64))
            top = pm < pin[0]
            want[top] = ozt[j, 0, i] * pm[top] / pin[0]
            assert np.allclose(o3[j, :nk, i], want, rtol=3e-6, atol=0), (j, i)
    assert (p[:, :nk, :] < pin[0]).any() and (p[:, :nk, :] > pin[-1]).any()


def cldfra3_case(ni=24, nj=6, nk=40, seed=21):
    from wrfchem_arc_interactions_b200 import synth
    dom = synth.make_domain(ni, nj, nk, seed=seed, cloudy_frac=0.4)
    rng = np.random.default_rng(seed)
    qv = dom["qv3d"].copy()
    qv[:, :, 5] = np.float32(1e-9)                                  # a dry column
    qv[:, 8:20, 6] *= np.float32(1.6)                               # a deep moist layer: fractional cloud decks
    qv[:, 3:30, 7] *= np.float32(1.25)
    return dom, np.ascontiguousarray(qv)


def test_cal_cldfra3_oracle(orc):
    """cal_cldfra3 (module_radiation_driver.F:3140-3274, find_cloudLayers / adjust_cloud* :3281-3599) restated in the oracle:
    resolved condensate -> 1; elsewhere 1 - sqrt((1 - RH) / (1 - RH_00)) capped at 0.9 with the grid-size and land / ocean
    dependent threshold, checked in closed form with the Flatau saturation polynomials; no fractional cloud in the two lowest
    levels; qc / qi only ever gain (sub-grid condensate of the cloud layers found), qs untouched; a dry column stays clear."""
    dom, qv = cldfra3_case()
    nk = dom["nk"]
    qc, qi, qs = dom["qc3d"].copy(), dom["qi3d"].copy(), dom["qs3d"].copy()
    cf = np.full(qc.shape, -3.0, np.float32)
    gridkm = 12.0
    orc.cal_cldfra3(dom["dims"], cf, qv, qc, qi, qs, dom["p3d"], dom["t3d"], dom["rho3d"], dom["xland"], gridkm)
    tile = (slice(None), slice(0, nk), slice(None))
    c = cf[tile]
    assert np.all(cf[:, nk, :] == -3.0) and np.array_equal(qs, dom["qs3d"])
    resolved = (dom["qc3d"][tile] > 1e-6) | (dom["qi3d"][tile] >= 1e-7) | (dom["qs3d"][tile] > 1e-5)
    assert np.all(c[resolved] == 1.0) and np.all((c == 0) | (c == 1) | ((c > 0) & (c <= np.float32(0.9))))
    frac = (c > 0) & (c < 1)
    assert frac.sum() > 50 and not frac[:, :2, :].any()            # kbot >= kts + 1: the two lowest levels never hold fractional cloud
    assert np.all(qc[tile] >= dom["qc3d"][tile]) and np.all(qi[tile] >= dom["qi3d"][tile])
    assert (qc[tile] > dom["qc3d"][tile]).any() and (qi[tile] > dom["qi3d"][tile]).any()
    grew = (qc[tile] > dom["qc3d"][tile]) | (qi[tile] > dom["qi3d"][tile])
    assert np.all(c[grew] > 0)                                      # condensate is only made up where there is cloud fraction
    assert not c[:, :, 5][~resolved[:, :, 5]].any()                 # the dry column

    def flatau(tk, ice):
        w = [.611583699E03, .444606896E02, .143177157E01, .264224321E-1, .299291081E-3, .203154182E-5, .702620698E-8, .379534310E-11, -.321582393E-13]
        i = [.609868993E03, .499320233E02, .184672631E01, .402737184E-1, .565392987E-3, .521693933E-5, .307839583E-7, .105785160E-9, .161444444E-12]
        x = max(-80.0, tk - 273.16)
        return sum(cc * x ** n for n, cc in enumerate(i if ice else w))
    assert abs(flatau(293.16, False) - 2339.0) < 5.0 and abs(flatau(253.16, True) - 103.2) < 0.5       # Pa, textbook values
    rh00 = {1.0: 0.7 + np.sqrt(1.0 / (25.0 + gridkm ** 3)), 2.0: 0.81 + np.sqrt(1.0 / (50.0 + gridkm ** 3))}
    checked = 0
    for jj, kk, ii in zip(*np.nonzero(frac & (dom["t3d"][tile] > 262.0))):
        tk, pp = float(dom["t3d"][jj, kk, ii]), float(dom["p3d"][jj, kk, ii])
        es = min(flatau(tk, False), 0.15 * pp)
        rh = min(max(float(qv[jj, kk, ii]) / (0.622 * es / (pp - es)), 0.01), 0.999)
        want = min(0.9, max(0.0, 1.0 - np.sqrt((1.0 - rh) / (1.0 - rh00[float(dom["xland"][jj, ii])]))))
        assert abs(float(c[jj, kk, ii]) - want) < 3e-4, (jj, kk, ii)
        checked += 1
    assert checked > 20


def tegen_case(ni=14, nj=6, nk=40, levsiz=12, seed=31):
    """Synthetic Tegen-style climatology: 12 months x 6 aerosol types on `levsiz` pressure levels (hPa, top down)."""
    from wrfchem_arc_interactions_b200 import synth
    rng = np.random.default_rng(seed)
    dom = synth.make_domain(ni, nj, nk, seed=seed)
    pin = np.linspace(20.0, 985.0, levsiz).astype(np.float32)                      # hPa
    prof = (1e-5 * np.exp(-(1000.0 - pin) / 250.0)).astype(np.float32)             # optical depth per Pa of layer thickness
    aerodm = (prof[None, None, None, :, None] * rng.uniform(0.2, 1.0, (6, 12, nj, levsiz, ni))).astype(np.float32)
    return dom, pin, np.ascontiguousarray(aerodm)


def test_aer_time_and_p_int_oracle(orc):
    """aer_time_int / aer_p_int (DRV:4236-4506): per aerosol type the time blend of ozn_time_int; then linear interpolation in
    pressure (model p in hPa) x the layer's interface-pressure difference, held below the data bottom, scaled by p / pin(1) above
    the data top; TOTAOD is the sum over types and levels."""
    dom, pin, aerodm = tegen_case()
    nk = dom["nk"]
    no_src, _, nj, levsiz, ni = aerodm.shape
    aerodt = np.zeros((no_src, nj, levsiz, ni), np.float32)
    orc.aer_time_int(dom["dims"], 0, 59.0, aerodm, aerodt, levsiz, 12, no_src)     # day 60: half way between mid-February and mid-March
    assert np.allclose(aerodt, 0.5 * (aerodm[:, 1] + aerodm[:, 2]), rtol=1e-6)
    aerod = np.full((no_src,) + dom["p3d"].shape, -1.0, np.float32); tot = np.full(dom["xland"].shape, -1.0, np.float32)
    orc.aer_p_int(dom["dims"], dom["p3d"], pin, levsiz, aerodt, aerod, no_src, dom["p8w"], tot)
    assert np.all(aerod[:, :, nk, :] == -1.0) and np.all(aerod[:, :, :nk, :] >= 0)
    for s in (0, 3, 5):
        for j in (0, nj - 1):
            for i in (0, ni // 2):
                pm = dom["p3d"][j, :nk, i].astype(np.float64) * 0.01
                want = np.interp(pm, pin.astype(np.float64), aerodt[s, j, :, i].astype(np.float64))
                top = pm < pin[0]
                want[top] = aerodt[s, j, 0, i] * pm[top] / pin[0]
                want *= dom["p8w"][j, :nk, i].astype(np.float64) - dom["p8w"][j, 1:nk + 1, i]
                assert np.allclose(aerod[s, j, :nk, i], want, rtol=5e-6, atol=0), (s, j, i)
    assert np.allclose(tot, aerod[:, :, :nk, :].astype(np.float64).sum(axis=(0, 2)), rtol=2e-6)
    assert 0.01 < tot.mean() < 5.0
