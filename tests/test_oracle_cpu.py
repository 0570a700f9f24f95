"""The oracle (CPU restatement of the reference's Fortran) against (a) the structural invariants the reference's code and
change notes state (SURVEY.md section 4 -- the reference ships no golden vectors or tests of its own), (b) committed golden
vectors (tests/golden/, regression pin of the oracle itself), (c) hypothesis-style properties of the index logic."""
import os

import numpy as np
import pytest

from conftest import ROOT, init, run_pair
from wrfchem_arc_interactions_b200 import abi, synth

CLN2D = ("swuptcln", "swdntcln", "swupbcln", "swdnbcln")


@pytest.fixture(scope="module")
def dom():
    return synth.make_domain(12, 6, 40, seed=11)


def test_clean_off_gives_zero_cln(orc, ktab, dom):
    """clean_atm_diag = 0 => every *CLN output is exactly 0 (SW:9473-9478, LW:11028-11033)."""
    init(orc, dom, ktab)
    sw = run_pair("sw", orc, dom, clean_atm_diag=0)
    lw = run_pair("lw", orc, dom, clean_atm_diag=0)
    for k in CLN2D + ("swupflxcln", "swdnflxcln"):
        assert not sw[k].any(), k
    for k in ("lwuptcln", "lwdntcln", "lwupbcln", "lwdnbcln", "lwupflxcln", "lwdnflxcln"):
        assert not lw[k].any(), k
    assert sw["swupt"].any() and lw["lwupt"].any()


def test_zero_aerosol_clean_equals_full_bitwise(orc, ktab):
    """AOD == 0 => clean == full bit for bit (same code, same inputs; the NCL layer relies on it,
    analysis_scripts/NCL_extraction_package/data_extraction_library.ncl:157-163)."""
    dom = synth.make_domain(12, 6, 40, seed=12, aerosol=False)
    init(orc, dom, ktab)
    sw = run_pair("sw", orc, dom)
    lw = run_pair("lw", orc, dom)
    for a, b in (("swupt", "swuptcln"), ("swdnb", "swdnbcln"), ("swupb", "swupbcln"), ("swupflx", "swupflxcln"), ("swdnflx", "swdnflxcln")):
        assert np.array_equal(sw[a], sw[b]), (a, b)
    for a, b in (("lwupt", "lwuptcln"), ("lwdnb", "lwdnbcln"), ("lwupflx", "lwupflxcln"), ("lwdnflx", "lwdnflxcln")):
        assert np.array_equal(lw[a], lw[b]), (a, b)


def test_night_columns(orc, ktab, dom):
    """coszen <= 0 => SW 2-D outputs 0, RTHRATENSW and GSW untouched, COSZR written everywhere (SW:10332, 11173-11199)."""
    from wrfchem_arc_interactions_b200 import radiation as R
    init(orc, dom, ktab)
    outs = R.alloc_outputs(dom, "sw")
    outs["rthratensw"][:] = 7.0
    outs["gsw"][:] = -3.0
    outs["swupt"][:] = 9.0
    sw = run_pair("sw", orc, dom, outs=outs)
    night = dom["xcoszen"] <= 0
    assert night.any() and (~night).any()
    assert np.all(sw["swupt"][night] == 0) and np.all(sw["swddir"][night] == 0) and np.all(sw["swcf"][night] == 0)
    assert np.all(sw["gsw"][night] == -3.0)
    assert np.all(sw["rthratensw"].transpose(0, 2, 1)[night] == 7.0)
    assert np.array_equal(sw["coszr"], dom["xcoszen"])
    assert np.all(sw["swupt"][~night] > 0)


def test_no_cloud_clear_equals_full(orc, ktab):
    """Cloud fraction == 0 => clear == full (streams only diverge at a cloudy layer: LW:3305-3311, SW zcloud = 0)."""
    dom = synth.make_domain(12, 6, 40, seed=13, cloudy_frac=0.0)
    init(orc, dom, ktab)
    sw, lw = run_pair("sw", orc, dom), run_pair("lw", orc, dom)
    assert np.array_equal(sw["swupt"], sw["swuptc"]) and np.array_equal(sw["swdnb"], sw["swdnbc"]) and not sw["swcf"].any()
    assert np.array_equal(lw["lwupt"], lw["lwuptc"]) and np.array_equal(lw["lwdnb"], lw["lwdnbc"]) and not lw["lwcf"].any()


def test_top_layer_heating_is_zero_and_energy_budget(orc, ktab, dom):
    """Heating rate of the top RRTMG layer is forced to 0 (SW:9433, LW:3407); TOA SW down = solcon/1368.22*sum(sflux)*mu0."""
    init(orc, dom, ktab)
    ncol, nlay = dom["ni"] * dom["nj"], dom["nk"] + 1
    dbg, taps = abi.alloc_debug(ncol, nlay, 112)
    sw = run_pair("sw", orc, dom, debug=dbg)
    sun = taps["laytrop"] >= 0
    assert np.all(taps["hr"][sun, nlay - 1] == 0)
    mu0 = dom["xcoszen"].ravel()[sun]
    toa = sw["swdnt"].ravel()[sun]
    s0 = float(dom["solcon"]) / 1368.22 * taps["sfluxzen"][sun].sum(axis=1)
    assert np.allclose(toa, s0 * mu0, rtol=2e-6)
    # absorbed + reflected <= incoming
    assert np.all(sw["swupt"].ravel()[sun] < toa) and np.all(sw["gsw"].ravel()[sun] > 0)


def test_sw_aerosol_column_cap(orc, ktab):
    """Per band sum(tau) > 6 => rescaled to 6 (SW:11034-11047): a 10x thicker aerosol changes nothing once capped."""
    d1 = synth.make_domain(6, 2, 40, seed=14, all_day=True)
    fac = 4000.0 / np.maximum(d1["tauaer400"].sum(axis=1, keepdims=True), 1e-9)
    for k in ("tauaer300", "tauaer400", "tauaer600", "tauaer999"):
        d1[k] = (d1[k] * fac).astype(np.float32)
    d2 = dict(d1)
    for k in ("tauaer300", "tauaer400", "tauaer600", "tauaer999"):
        d2[k] = (d1[k] * 10.0).astype(np.float32)
    init(orc, d1, ktab)
    a, b = run_pair("sw", orc, d1), run_pair("sw", orc, d2)
    assert np.allclose(a["swdnb"], b["swdnb"], rtol=2e-5, atol=1e-3)
    assert np.all(a["swdnb"] < a["swdnbcln"])


def test_negative_aod_is_an_error(orc, ktab):
    from wrfchem_arc_interactions_b200 import radiation as R
    dom = synth.make_domain(4, 2, 40, seed=15, all_day=True)
    for k in ("tauaer300", "tauaer400", "tauaer999"):
        dom[k] = dom[k].copy()
    dom["tauaer400"][0, :, 0] = -1.0
    init(orc, dom, ktab)
    with pytest.raises(R.RadiationError) as e:
        run_pair("sw", orc, dom)
    assert e.value.code == 5 and "Negative total optical depth" in str(e.value)


def test_missing_aerosol_field_is_an_error(orc, ktab):
    from wrfchem_arc_interactions_b200 import radiation as R
    dom = synth.make_domain(4, 2, 40, seed=15)
    del dom["tauaer600"]
    init(orc, dom, ktab)
    with pytest.raises(R.RadiationError) as e:
        run_pair("sw", orc, dom)
    assert e.value.code == 4


def test_index_ranges_and_mask_statistics(orc, ktab, dom):
    init(orc, dom, ktab)
    ncol, nlay = dom["ni"] * dom["nj"], dom["nk"] + 1
    dbg, t = abi.alloc_debug(ncol, nlay, 112)
    run_pair("sw", orc, dom, debug=dbg)
    sun = t["laytrop"] >= 0
    assert t["jp"][sun].min() >= 1 and t["jp"][sun].max() <= 58
    for k in ("jt", "jt1"):
        assert t[k][sun].min() >= 1 and t[k][sun].max() <= 4
    assert set(np.unique(t["indfor"][sun])) <= {1, 2, 3} and t["indself"][sun].max() <= 9
    # laytrop = number of layers with ln(p) > 4.56, i.e. p > 95.58 hPa
    p = dom["p3d"].transpose(0, 2, 1).reshape(ncol, -1)[:, : dom["nk"]] / 100.0
    assert np.array_equal(t["laytrop"][sun], (np.log(p[sun].astype(np.float32)) > np.float32(4.56)).sum(axis=1))
    # McICA: cloud-free layers have no cloudy sub-column; overcast layers are cloudy in every sub-column
    cf = dom["cldfra3d"].transpose(0, 2, 1).reshape(ncol, -1)[:, : dom["nk"]]
    m = t["cldmask"][:, : dom["nk"], :]
    assert not m[sun][cf[sun] == 0].any() and m[sun][cf[sun] == 1].all()
    part = (cf > 0.2) & (cf < 0.8) & sun[:, None]
    if part.any():
        frac = m[part].mean(axis=1)
        assert abs(frac.mean() - cf[part].mean()) < 0.08


def test_golden_vectors(orc, ktab):
    """Regression pin of the oracle: fixtures made by tests/golden/make_golden.py from this same oracle at commit time."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "oracle_c1_slice.npz"))
    dom = synth.make_domain(int(g["ni"]), int(g["nj"]), int(g["nk"]), seed=int(g["seed"]))
    init(orc, dom, ktab)
    sw, lw = run_pair("sw", orc, dom), run_pair("lw", orc, dom)
    for k in g.files:
        if k.startswith("sw_"):
            assert np.allclose(sw[k[3:]], g[k], rtol=1e-6, atol=1e-6), k
        elif k.startswith("lw_"):
            assert np.allclose(lw[k[3:]], g[k], rtol=1e-6, atol=1e-6), k


def test_kissvec_known_answers():
    """kissvec (SW:1900-1932) restated independently in Python integers: first draws from fixed seeds."""
    import ctypes as C
    def kiss(s):
        m32 = 0xFFFFFFFF
        s[0] = (69069 * s[0] + 1327217885) & m32
        x = s[1]; x ^= (x << 13) & m32; x ^= x >> 17; x ^= (x << 5) & m32; s[1] = x
        s[2] = (18000 * (s[2] & 65535) + (s[2] >> 16)) & m32
        s[3] = (30903 * (s[3] & 65535) + (s[3] >> 16)) & m32
        k = (s[0] + s[1] + ((s[2] << 16) & m32) + s[3]) & m32
        k = k - (1 << 32) if k >= (1 << 31) else k
        return np.float32(np.float32(np.float32(k) * np.float32(2.328306e-10)) + np.float32(0.5))
    s = [123456789, 362436069, 521288629, 916191069]
    draws = [kiss(s) for _ in range(1000)]
    assert 0.0 <= min(draws) and max(draws) <= 1.0 and abs(float(np.mean(draws)) - 0.5) < 0.03
    # the generator is a pure function of its state: same seeds, same stream
    s2 = [123456789, 362436069, 521288629, 916191069]
    assert [kiss(s2) for _ in range(5)] == draws[:5]


def inline_table(name):
    """One table of data/rrtmg_inline_tables.bin (layout: tools/extract_inline_tables.py), flat, Fortran element order."""
    import struct
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "wrfchem-arc-interactions_b200", "data", "rrtmg_inline_tables.bin")
    blob = open(path, "rb").read()
    assert blob[:8] == b"ARCTBL1\0"
    n = struct.unpack_from("<I", blob, 8)[0]
    for e in range(n):
        off = 12 + 76 * e
        if blob[off:off + 32].split(b"\0")[0].decode() == name:
            ndim = struct.unpack_from("<I", blob, off + 32)[0]
            dims = struct.unpack_from("<4I", blob, off + 36)[:ndim]
            return np.frombuffer(blob, "<f4", int(np.prod(dims)) if ndim else 1, struct.unpack_from("<Q", blob, off + 68)[0]).copy()
    raise KeyError(name)


def tegen_aerod(dom, seed=5, types=(0, 1, 2, 3, 4, 5)):
    """AEROD(i,k,j,1:6) in C order (6, nj, nkm, ni): layer optical depths at 0.55 um of the ECMWF aerosol types."""
    rng = np.random.default_rng(seed)
    nj, nkm, ni = dom["t3d"].shape
    prof = np.exp(-np.arange(nkm) / 6.0)[None, :, None]
    a = np.zeros((6, nj, nkm, ni), np.float32)
    for t in types:
        a[t] = (rng.uniform(0.0, 0.03, (nj, 1, ni)) * prof).astype(np.float32)
    return a


def test_aer_opt_1_ecmwf_aerosol_types(orc, ktab):
    """aer_opt = 1 -> iaer = 6 (SW:9201-9205, 9313-9341, 11083-11100): zero AEROD is the aerosol-free atmosphere bit for bit; one
    aerosol type alone equals the direct specification (the tauaer3d_sw input) of rsrtaua * AEROD, rsrpiza, rsrasya per band;
    more aerosol dims the surface; the clean diagnostic, undefined in the reference for this option, is refused."""
    from wrfchem_arc_interactions_b200 import radiation as R
    dom = synth.make_domain(8, 3, 40, seed=16, all_day=True)
    init(orc, dom, ktab)
    nj, nkm, ni = dom["t3d"].shape
    base = run_pair("sw", orc, dom, clean_atm_diag=0, aer_ra_feedback=0)
    zero = run_pair("sw", orc, dom, clean_atm_diag=0, aer_ra_feedback=0, aer_opt=1, no_src=6, aerod=np.zeros((6, nj, nkm, ni), np.float32))
    for k in base:
        assert np.array_equal(base[k], zero[k]), k
    aerod = tegen_aerod(dom)
    mix = run_pair("sw", orc, dom, clean_atm_diag=0, aer_ra_feedback=0, aer_opt=1, no_src=6, aerod=aerod)
    assert np.all(mix["swdnb"] < base["swdnb"]) and np.all(mix["swdnbc"] < base["swdnbc"])
    one = tegen_aerod(dom, types=(1,))
    rsr = {n: inline_table("sw_" + n).reshape(6, 14) for n in ("rsrtaua", "rsrpiza", "rsrasya")}      # [type][band]
    tau = np.stack([rsr["rsrtaua"][1, b] * one[1] for b in range(14)]).astype(np.float32)
    ssa = np.stack([np.full_like(one[1], rsr["rsrpiza"][1, b]) for b in range(14)]); asy = np.stack([np.full_like(one[1], rsr["rsrasya"][1, b]) for b in range(14)])
    a = run_pair("sw", orc, dom, clean_atm_diag=0, aer_ra_feedback=0, aer_opt=1, no_src=6, aerod=one)
    b = run_pair("sw", orc, dom, clean_atm_diag=0, aer_ra_feedback=0, aer_opt=2, tauaer3d_sw=tau, ssaaer3d_sw=ssa, asyaer3d_sw=asy)
    for k in ("swdnb", "swupt", "swdnbc", "swddir", "gsw"):
        assert np.allclose(a[k], b[k], rtol=2e-5, atol=2e-3), k
    with pytest.raises(R.RadiationError) as e:
        run_pair("sw", orc, dom, clean_atm_diag=1, aer_opt=1, no_src=6, aerod=aerod)
    assert "undefined" in str(e.value)
    with pytest.raises(R.RadiationError):
        run_pair("sw", orc, dom, clean_atm_diag=0, aer_ra_feedback=0, aer_opt=1, no_src=6)
