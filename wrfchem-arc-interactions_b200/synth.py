"""Seeded synthetic WRF-layout column sets for parity tests and the benchmark.

Everything is float32 in WRF memory order: 3-D fields are numpy arrays of shape
(nj, nk, ni) in C order == Fortran (i,k,j) with i contiguous; 2-D fields (nj, ni);
flux profiles (nj, nk+2, ni).  Memory bounds equal tile bounds unless `halo` > 0, in
which case ims = its-halo etc. and the halo is filled with NaN (the boundary must never
read or write it; module_radiation_driver.F tile/memory dims convention).
"""
from __future__ import annotations

import numpy as np

SEED = 20120650
LW_BAND_CENTRES_CM = np.array([180., 425., 565., 665., 760., 900., 1030., 1130., 1285., 1435., 1640., 1940., 2165.,
                               2315., 2490., 2925.])


def _qsat(T, p_pa):
    es = 611.2 * np.exp(17.67 * (T - 273.15) / (T - 29.65))
    es = np.minimum(es, 0.5 * p_pa)
    return 0.622 * es / (p_pa - 0.378 * es)


def make_domain(ni, nj, nk, seed=SEED, p_top=5000.0, cloudy_frac=0.4, night_frac=0.25, with_re=False,
                aerosol=True, halo=0, all_day=False):
    """Return dict(dims=..., fields...) for an ni x nj x nk tile."""
    rng = np.random.default_rng(seed)
    f32 = np.float32
    ncol = ni * nj
    psfc = rng.uniform(95000.0, 103500.0, ncol)
    k = np.arange(nk + 1)
    sigma = (1.0 - k / nk) ** 1.35
    sigma = sigma * (1.0 + 0.15 * (k / nk) * (1 - k / nk))
    sigma[0], sigma[-1] = 1.0, 0.0
    p8w = (p_top + sigma[None, :] * (psfc[:, None] - p_top)).astype(f32)         # (ncol, nk+1) Pa
    p3d = (0.5 * (p8w[:, :-1].astype(np.float64) + p8w[:, 1:])).astype(f32)
    tsk = rng.uniform(255.0, 310.0, ncol)
    ttrop = rng.uniform(200.0, 220.0, ncol)

    def temp(p):
        z = 7.5 * np.log(psfc[:, None] / p)                                         # km
        t = tsk[:, None] - 2.0 - 6.5 * z
        ztrop = (tsk[:, None] - 2.0 - ttrop[:, None]) / 6.5
        above = z > ztrop
        t = np.where(above, ttrop[:, None] + 1.5 * (z - ztrop), t)
        return t
    t3d = temp(p3d.astype(np.float64)) + rng.uniform(-2.0, 2.0, (ncol, nk))
    t8w = temp(np.maximum(p8w.astype(np.float64), 1.0))
    t8w[:, 1:-1] = 0.5 * (t3d[:, :-1] + t3d[:, 1:])
    t8w[:, 0] = tsk - 1.0
    t8w[:, -1] = t3d[:, -1] + 0.5 * (t3d[:, -1] - t3d[:, -2])
    rh = rng.uniform(0.2, 0.95, (ncol, nk))
    qv = rh * _qsat(t3d, p3d.astype(np.float64))
    qv = np.where(p3d > 30000.0, qv, 3.0e-6)
    qv = np.maximum(qv, 1.0e-12)
    rho = p3d / (287.0 * t3d * (1.0 + 0.61 * qv))
    dz = (p8w[:, :-1].astype(np.float64) - p8w[:, 1:]) / (rho * 9.81)
    pi3d = (p3d.astype(np.float64) / 1.0e5) ** (2.0 / 7.0)

    # clouds
    cld = np.zeros((ncol, nk)); qc = np.zeros((ncol, nk)); qi = np.zeros((ncol, nk)); qs = np.zeros((ncol, nk))
    cloudy_cols = rng.random(ncol) < cloudy_frac
    ndeck = rng.integers(1, 4, ncol)
    kk = np.arange(nk)[None, :]
    for deck in range(3):
        base = rng.integers(1, max(2, int(0.75 * nk)), ncol)[:, None]
        thick = rng.integers(1, 5, ncol)[:, None]
        cf = np.where(rng.random(ncol) < 0.3, 1.0, rng.uniform(0.05, 1.0, ncol))[:, None]
        m = cloudy_cols[:, None] & (deck < ndeck)[:, None] & (kk >= base) & (kk < base + thick) & (kk < nk - 2)
        cld = np.where(m, np.maximum(cld, cf), cld)
        liq = m & (p3d > 60000.0) & (t3d > 253.0)
        qc = np.where(liq, 10.0 ** rng.uniform(-5.0, -3.3, (ncol, nk)), qc)
        ice = m & (t3d < 263.0)
        qi = np.where(ice, 10.0 ** rng.uniform(-6.0, -4.0, (ncol, nk)), qi)
        qs = np.where(ice, 0.3 * qi, qs)
        neither = m & (qc == 0.0) & (qi == 0.0)
        qc = np.where(neither, 10.0 ** rng.uniform(-5.0, -4.0, (ncol, nk)), qc)
    albedo = np.where(rng.random(ncol) < 0.05, 0.8, rng.uniform(0.05, 0.35, ncol))
    emiss = rng.uniform(0.9, 1.0, ncol)
    xland = np.where(rng.random(ncol) < 0.5, 1.0, 2.0)
    coszen = rng.uniform(0.05, 1.0, ncol)
    if not all_day:
        night = rng.random(ncol) < night_frac
        coszen = np.where(night, -rng.uniform(0.0, 0.5, ncol), coszen)

    # aerosol optics at the four chem wavelengths + 16 LW bands
    aod400 = np.minimum(np.exp(rng.normal(np.log(0.25), 0.8, ncol)), 4.5)
    zmid = 7.5 * np.log(psfc[:, None] / p3d)
    prof = np.exp(-zmid / 1.5) * dz
    tau400 = aod400[:, None] * prof / prof.sum(axis=1, keepdims=True)
    ang = rng.uniform(0.3, 2.0, (ncol, 1))
    w0 = rng.uniform(0.8, 0.99, (ncol, nk)); g0 = rng.uniform(0.55, 0.75, (ncol, nk))
    aer = {}
    for wl in (300, 400, 600, 999):
        aer["tauaer%d" % wl] = tau400 * (400.0 / wl) ** ang
        aer["waer%d" % wl] = np.clip(w0 - 0.02 * (wl - 400) / 600.0, 0.0, 1.0)
        aer["gaer%d" % wl] = np.clip(g0 - 0.05 * (wl - 400) / 600.0, 0.0, 1.0)
    lam = 1.0e4 / LW_BAND_CENTRES_CM
    for b in range(16):
        aer["tauaerlw%d" % (b + 1)] = tau400 * (0.4 / lam[b]) ** ang * (1.0 - w0)
    if not aerosol:
        for kname in aer:
            aer[kname] = np.zeros_like(aer[kname])

    h = int(halo)
    mi, mj = ni + 2 * h, nj + 2 * h

    def to3(a, nlev=nk):
        out = np.full((mj, nlev, mi), np.nan if h else 0.0, dtype=f32)
        out[h:h + nj, :, h:h + ni] = np.asarray(a, dtype=np.float64).reshape(nj, ni, nlev).transpose(0, 2, 1).astype(f32)
        return np.ascontiguousarray(out)

    def to2(a):
        out = np.full((mj, mi), np.nan if h else 0.0, dtype=f32)
        out[h:h + nj, h:h + ni] = np.asarray(a, dtype=np.float64).reshape(nj, ni).astype(f32)
        return np.ascontiguousarray(out)

    # 3-D arrays are dimensioned kms:kme with kme = nk+1 (WRF: kme = kde = kte+1)
    def to3k(a):
        b = np.concatenate([np.asarray(a, dtype=np.float64), np.zeros((ncol, 1))], axis=1)
        return to3(b, nk + 1)

    d = dict(
        dims=dict(ids=1, ide=ni + 1, jds=1, jde=nj + 1, kds=1, kde=nk + 1,
                  ims=1 - h, ime=ni + h, jms=1 - h, jme=nj + h, kms=1, kme=nk + 1,
                  its=1, ite=ni, jts=1, jte=nj, kts=1, kte=nk),
        ni=ni, nj=nj, nk=nk, halo=h, p_top=float(p_top),
        t3d=to3k(t3d), t8w=to3(t8w, nk + 1), p3d=to3k(p3d), p8w=to3(p8w, nk + 1), pi3d=to3k(pi3d), rho3d=to3k(rho),
        dz8w=to3k(dz), cldfra3d=to3k(cld), qv3d=to3k(qv), qc3d=to3k(qc), qr3d=to3k(np.zeros((ncol, nk))),
        qi3d=to3k(qi), qs3d=to3k(qs), qg3d=to3k(np.zeros((ncol, nk))),
        xcoszen=to2(coszen), albedo=to2(albedo), tsk=to2(tsk), xland=to2(xland), xice=to2(np.zeros(ncol)),
        snow=to2(np.zeros(ncol)), emiss=to2(emiss),
        solcon=np.float32(1370.0 * 1.01), r=np.float32(287.0), g=np.float32(9.81),
    )
    if with_re:
        d["re_cloud"] = to3k(np.where(qc > 0, rng.uniform(4e-6, 20e-6, (ncol, nk)), 0.0))
        d["re_ice"] = to3k(np.where(qi > 0, rng.uniform(10e-6, 80e-6, (ncol, nk)), 0.0))
        d["re_snow"] = to3k(np.where(qs > 0, rng.uniform(50e-6, 300e-6, (ncol, nk)), 0.0))
    for kname, v in aer.items():
        d[kname] = to3k(v)
    # pi3d/p3d in the unused top memory level must stay finite for safety
    return d


MOSAIC_SPECIES = ("so4", "no3", "cl", "nh4", "na", "oin", "oc", "bc", "water")
MODAL_SPECIES = {  # MADE/SORGAM names per mode (registry.chem:3820): i = Aitken, j = accumulation, c = coarse
    "i": ("so4_so4ai", "nh4_nh4ai", "no3_no3ai", "na_naai", "cl_clai", "oc_orgaro1i", "oc_orgalk1i", "oc_orgpai", "bc_eci", "oin_p25i", "water_h2oai"),
    "j": ("so4_so4aj", "nh4_nh4aj", "no3_no3aj", "na_naaj", "cl_claj", "oc_orgaro1j", "oc_orgalk1j", "oc_orgpaj", "bc_ecj", "oin_p25j", "water_h2oaj"),
    "c": ("oin_antha", "na_seas", "oin_soila"),
}
MODAL_SIGMAG = (1.7, 2.0, 2.5)


def make_aerosol(dom, nbin=8, modal=False, seed=SEED + 7):
    """Synthetic aerosol mass / number fields in WRF layout for the optics stage.  Returns (bins, alt, sigmag):
    sectional: `nbin` dicts {species: ug/kg array, "num": #/kg}; modal: three dicts (Aitken, accumulation, coarse)."""
    rng = np.random.default_rng(seed)
    nj, nkm, ni = dom["t3d"].shape
    f32 = np.float32
    rho = np.where(dom["rho3d"] > 0, dom["rho3d"], 1.0).astype(np.float64)
    alt = (1.0 / rho).astype(f32)
    alt[~np.isfinite(alt)] = 1.0
    z = np.cumsum(np.nan_to_num(dom["dz8w"].astype(np.float64)), axis=1) / 1000.0           # km
    total = 12.0 * np.exp(-z / 2.0) * np.exp(rng.normal(0.0, 0.5, (nj, 1, ni)))            # ug/kg dry aerosol, all sizes
    dens = dict(so4=1.8, no3=1.8, cl=2.2, nh4=1.8, na=2.2, oin=2.6, oc=1.0, bc=1.7, water=1.0)
    comp = dict(so4=0.30, no3=0.10, cl=0.02, nh4=0.12, na=0.03, oin=0.18, oc=0.20, bc=0.05)
    bins = []
    if not modal:
        lo, hi = 3.90625e-6, 1.0e-3
        edges = lo * (hi / lo) ** (np.arange(nbin + 1) / nbin)
        dc = np.sqrt(edges[:-1] * edges[1:])                                                # cm
        wmass = np.exp(-0.5 * (np.log(dc / 3.0e-5) / np.log(2.2)) ** 2) + 0.25 * np.exp(-0.5 * (np.log(dc / 3.0e-4) / np.log(1.8)) ** 2)
        wmass /= wmass.sum()
        for b in range(nbin):
            spec = {}
            vol = 0.0
            for sp, fr in comp.items():
                tilt = 1.0 + (0.6 if sp in ("oin", "na", "cl") else -0.3) * (b - nbin / 2) / nbin
                m = total * wmass[b] * fr * tilt * rng.uniform(0.8, 1.2, (nj, nkm, ni))
                spec[sp] = m.astype(f32)
                vol = vol + m * 1e-6 / dens[sp]                                             # dry volume, cm3 per kg-air
            water = 0.6 * sum(spec[s_].astype(np.float64) for s_ in ("so4", "no3", "nh4", "na", "cl")) * rng.uniform(0.2, 1.5, (nj, 1, ni))
            spec["water"] = water.astype(f32)
            dmean = dc[b] * rng.uniform(0.85, 1.15, (nj, nkm, ni))
            spec["num"] = (vol / (np.pi / 6.0 * dmean ** 3)).astype(f32)                    # #/kg-air
            bins.append(spec)
        return bins, alt, None
    shares = {"i": 0.05, "j": 0.70, "c": 0.25}
    dgs = {"i": 3.0e-6, "j": 1.5e-5, "c": 1.0e-4}                                           # number median diameters, cm
    for mi, md in enumerate(("i", "j", "c")):
        spec = {}
        vol = 0.0
        names = MODAL_SPECIES[md]
        for nm in names:
            cls = nm.split("_")[0]
            if cls == "water":
                m = 0.5 * total * shares[md] * rng.uniform(0.2, 1.0, (nj, 1, ni))
            else:
                m = total * shares[md] * comp.get(cls, 0.1) / max(1, sum(1 for q in names if q.split("_")[0] == cls)) * rng.uniform(0.8, 1.2, (nj, nkm, ni))
                vol = vol + m * 1e-6 / dens[cls]
            spec[nm] = m.astype(f32)
        ls = np.log(MODAL_SIGMAG[mi])
        dg = dgs[md] * rng.uniform(0.8, 1.25, (nj, nkm, ni))
        spec["num"] = (vol / (np.pi / 6.0 * dg ** 3 * np.exp(4.5 * ls * ls))).astype(f32)
        bins.append(spec)
    return bins, alt, MODAL_SIGMAG


CONFIGS = {
    # name: (ni, nj, nk, kwargs)   -- BASELINE.md section 3
    "C1": (32, 32, 40, dict()),
    "C2": (425, 300, 50, dict()),
    "C3": (2000, 2000, 60, dict()),
    "C4": (425, 300, 50, dict(cloudy_frac=1.0, with_re=True)),
    "C5": (1000, 1000, 100, dict()),
}
