"""Host-side mirror of the reference's operator interface for the radiation hot path.

Same entry-point names and argument meaning as the Fortran the reference patches:

    rrtmg_swinit / rrtmg_lwinit   module_ra_rrtmg_sw.F:11211, module_ra_rrtmg_lw.F:12845
    RRTMG_SWRAD                   module_ra_rrtmg_sw.F:9901
    RRTMG_LWRAD                   module_ra_rrtmg_lw.F:11451

Arguments are keyword arguments named exactly like the Fortran dummies; arrays are numpy
float32 arrays in WRF memory order (host) or torch CUDA tensors (device, zero-copy).
A non-zero status raises RadiationError (the reference calls wrf_error_fatal / stop).

All arithmetic happens in libarcrad.so (csrc/, hand-written CUDA for sm_100a).  There is
no CPU fallback: if the library is missing this module fails loudly at import of the
library handle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ARC_RAD_LIB") or os.path.join(_HERE, "csrc", "libarcrad.so")
INLINE_TABLES = os.path.join(_HERE, "data", "rrtmg_inline_tables.bin")


class RadiationError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("arc_rad status %d: %s" % (code, msg))
        self.code = code


def _is_device(a):
    return hasattr(a, "data_ptr") and getattr(a, "is_cuda", False)


def _sync_producers(arrays):
    """Stream contract of the C ABI for device arrays (include/arc_rad.h, "Streams"): the library works on its own
    non-blocking streams and does not order itself against the caller's, so everything that produces the arrays must have
    finished before the call.  The arrays here are torch tensors: wait for the current torch stream of their device."""
    for a in arrays:
        if _is_device(a):
            import torch
            torch.cuda.current_stream(a.device).synchronize()
            return


def _ptr(a):
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return abi.fptr(int(a.data_ptr()))
    return abi.fptr(a)


class RadLib:
    """Thin binding of a shared library exporting the arc_rad C ABI (prefix selects the symbol family)."""

    def __init__(self, path, prefix="arc_rad_"):
        if not os.path.exists(path):
            raise ImportError("native library %s not found: run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
        self.lib = C.CDLL(path)
        self.prefix = prefix
        L = self.lib
        g = lambda n: getattr(L, prefix + n)
        self._init = g("init"); self._init.restype = C.c_int
        self._init.argtypes = [C.POINTER(abi.ArcConfig), C.c_char_p, C.c_char_p]
        self._err = g("last_error"); self._err.restype = C.c_char_p
        self._nlay = g("lw_nlayers"); self._nlay.restype = C.c_int
        if prefix == "arc_rad_":
            self._sw = L.arc_rad_sw_debug; self._lw = L.arc_rad_lw_debug
        else:
            self._sw = g("sw"); self._lw = g("lw")
        self._sw.restype = C.c_int; self._lw.restype = C.c_int
        self._sw.argtypes = [C.POINTER(abi.ArcDims), C.POINTER(abi.ArcSwIn), C.POINTER(abi.ArcSwOut), C.POINTER(abi.ArcDebug)]
        self._lw.argtypes = [C.POINTER(abi.ArcDims), C.POINTER(abi.ArcLwIn), C.POINTER(abi.ArcLwOut), C.POINTER(abi.ArcDebug)]
        self.initialised = False

    def last_error(self):
        return (self._err() or b"").decode(errors="replace")

    def check(self, rc):
        if rc != 0:
            raise RadiationError(rc, self.last_error())

    # rrtmg_swinit + rrtmg_lwinit
    def init(self, p_top, kme, sw_data, lw_data, cp=1004.5, device=-1, inline_tables=None):
        cfg = abi.ArcConfig(cp=float(cp), p_top=float(p_top), kme=int(kme), device=int(device),
                            inline_tables=(inline_tables or INLINE_TABLES).encode())
        self.check(self._init(C.byref(cfg), sw_data.encode(), lw_data.encode()))
        self.initialised = True
        return self

    def lw_nlayers(self):
        return int(self._nlay())

    @staticmethod
    def _fill(struct, names, kw, used):
        for n in names:
            v = kw.get(n)
            used.add(n)
            setattr(struct, n, _ptr(v))

    def RRTMG_SWRAD(self, dims, debug=None, **kw):
        """kw: Fortran dummy names -> arrays / scalars.  F_Qx flags: True/False or omitted (= not PRESENT)."""
        d = abi.make_dims(dims) if isinstance(dims, dict) else dims
        si, so = self._build_sw(kw)
        _sync_producers(kw.values())
        self.check(self._sw(C.byref(d), C.byref(si), C.byref(so), C.byref(debug) if debug is not None else None))

    def _build_sw(self, kw):
        si, so = abi.ArcSwIn(), abi.ArcSwOut()
        used = set()
        dev = [_is_device(v) for v in kw.values() if v is not None and (hasattr(v, "data_ptr") or isinstance(v, np.ndarray))]
        if dev and any(dev) and not all(dev):
            raise ValueError("mix of host and device arrays")
        si.memspace = abi.ARC_MEM_DEVICE if (dev and dev[0]) else abi.ARC_MEM_HOST
        si.variant_mask = int(kw.get("variant_mask", 0)); used.add("variant_mask")
        for n in abi.SW_IN_SCALARS_F:
            setattr(si, n, float(kw.get(n, 0.0))); used.add(n)
        for n in abi.SW_IN_SCALARS_I:
            dflt = -1 if n.startswith("f_q") else 0
            v = kw.get(n, dflt)
            setattr(si, n, int(v) if v is not None else dflt); used.add(n)
        self._fill(si, abi.SW_IN_3D + abi.SW_IN_2D, kw, used)
        self._fill(so, abi.SW_OUT_3D + abi.SW_OUT_2D + abi.SW_OUT_PROF + abi.SW_OUT_EXT, kw, used)
        unknown = set(kw) - used
        if unknown:
            raise TypeError("RRTMG_SWRAD: unknown arguments %s" % sorted(unknown))
        for req in ("rthratensw", "gsw", "swcf", "coszr", "swddir", "swddni", "swddif", "xcoszen", "albedo", "t3d", "t8w",
                    "p3d", "p8w", "pi3d", "qv3d", "tsk", "xland", "xice", "snow"):
            if kw.get(req) is None:
                raise TypeError("RRTMG_SWRAD: required argument %s missing" % req)
        return si, so

    def RRTMG_LWRAD(self, dims, debug=None, **kw):
        d = abi.make_dims(dims) if isinstance(dims, dict) else dims
        li, lo = self._build_lw(kw)
        _sync_producers(kw.values())
        self.check(self._lw(C.byref(d), C.byref(li), C.byref(lo), C.byref(debug) if debug is not None else None))

    def RRTMG_LWSW(self, dims, lw_kw, sw_kw):
        """One radiation step, RRTMG_LWRAD then RRTMG_SWRAD (the order of radiation_driver), through arc_rad_lwsw: with host
        arrays both adapters share one upload / compute / download pipeline."""
        d = abi.make_dims(dims) if isinstance(dims, dict) else dims
        li, lo = self._build_lw(lw_kw)
        si, so = self._build_sw(sw_kw)
        fn = self.lib.arc_rad_lwsw
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(abi.ArcDims), C.POINTER(abi.ArcLwIn), C.POINTER(abi.ArcLwOut), C.POINTER(abi.ArcSwIn), C.POINTER(abi.ArcSwOut)]
        _sync_producers(list(lw_kw.values()) + list(sw_kw.values()))
        self.check(fn(C.byref(d), C.byref(li), C.byref(lo), C.byref(si), C.byref(so)))

    def _build_lw(self, kw):
        li, lo = abi.ArcLwIn(), abi.ArcLwOut()
        used = set()
        dev = [_is_device(v) for v in kw.values() if v is not None and (hasattr(v, "data_ptr") or isinstance(v, np.ndarray))]
        if dev and any(dev) and not all(dev):
            raise ValueError("mix of host and device arrays")
        li.memspace = abi.ARC_MEM_DEVICE if (dev and dev[0]) else abi.ARC_MEM_HOST
        li.variant_mask = int(kw.get("variant_mask", 0)); used.add("variant_mask")
        for n in abi.LW_IN_SCALARS_F:
            setattr(li, n, float(kw.get(n, 0.0))); used.add(n)
        for n in abi.LW_IN_SCALARS_I:
            dflt = -1 if n.startswith("f_q") else 0
            v = kw.get(n, dflt)
            setattr(li, n, int(v) if v is not None else dflt); used.add(n)
        self._fill(li, abi.LW_IN_3D + abi.LW_IN_2D, kw, used)
        for b in range(16):
            n = "tauaerlw%d" % (b + 1)
            li.tauaerlw[b] = _ptr(kw.get(n)); used.add(n)
        self._fill(lo, abi.LW_OUT_3D + abi.LW_OUT_2D + abi.LW_OUT_PROF + abi.LW_OUT_EXT, kw, used)
        unknown = set(kw) - used
        if unknown:
            raise TypeError("RRTMG_LWRAD: unknown arguments %s" % sorted(unknown))
        for req in ("rthratenlw", "glw", "olr", "lwcf", "emiss", "t3d", "t8w", "p3d", "p8w", "pi3d", "qv3d", "tsk",
                    "xland", "xice", "snow"):
            if kw.get(req) is None:
                raise TypeError("RRTMG_LWRAD: required argument %s missing" % req)
        return li, lo


    # optical_averaging (WRF-Chem chem/module_optical_averaging.F; restated)
    def domain_statistics(self, dims, fields, names=None, morans=True, percentiles=True, trim=0, region=None):
        """The 13 statistics of calc_standard_stats (misc_stats_library.ncl:396-461) for 2-D fields, reduced on the device:
        avg, stddev (N-1), min, max, median, lower_quartile, upper_quartile, p05, p95 (with `percentiles`), standard_error,
        morans_i and corrected_standard_error = SE * I (with `morans`), N.  `trim` cells are cut from every edge of the tile
        first, as calculate_domain_stats does with domain_trim@trim = 5 (data_extraction_library.ncl:318-322); `region` =
        (lon_start, lon_end, lat_start, lat_end), the 0-based inclusive index box of its region_select branch (ncl:291-316, from
        wrf_user_ll_to_ij), takes precedence over `trim` as in the reference.
        `fields`: list of numpy arrays (host) or torch CUDA tensors (device); returns decomposition.stats_from_sums(...)."""
        from . import decomposition
        d = abi.make_dims(dims) if isinstance(dims, dict) else abi.ArcDims.from_buffer_copy(dims)
        if region is not None:
            i0, i1, j0, j1 = (int(x) for x in region)
            if not (0 <= i0 <= i1 <= d.ite - d.its and 0 <= j0 <= j1 <= d.jte - d.jts):
                raise ValueError("region outside the tile")
            d.ite = d.its + i1; d.its = d.its + i0; d.jte = d.jts + j1; d.jts = d.jts + j0
        elif trim:
            d.its += trim; d.ite -= trim; d.jts += trim; d.jte -= trim
        dev = [_is_device(f) for f in fields]
        if any(dev) and not all(dev):
            raise ValueError("mix of host and device arrays")
        ms = abi.ARC_MEM_DEVICE if dev[0] else abi.ARC_MEM_HOST
        n = len(fields)
        ptrs = (abi.c_fp * n)(*[abi.fptr(int(f.data_ptr())) if dev[0] else abi.fptr(f) for f in fields])
        L = self.lib
        L.arc_rad_domain_stats.restype = C.c_int
        L.arc_rad_domain_stats.argtypes = [C.POINTER(abi.ArcDims), C.c_int, C.c_int, C.POINTER(abi.c_fp), C.c_void_p]
        L.arc_rad_morans_i.restype = C.c_int
        L.arc_rad_morans_i.argtypes = [C.POINTER(abi.ArcDims), C.c_int, C.c_int, C.POINTER(abi.c_fp), C.c_void_p]
        L.arc_rad_percentiles.restype = C.c_int
        L.arc_rad_percentiles.argtypes = [C.POINTER(abi.ArcDims), C.c_int, C.c_int, C.POINTER(abi.c_fp), C.c_int, abi.c_fp, C.c_void_p]
        perc = np.array([50.0, 25.0, 75.0, 5.0, 95.0], np.float32)          # median, quartiles, 5th / 95th (ncl:441-445)
        if dev[0]:
            import torch
            st = torch.zeros(n, 5, dtype=torch.float64, device=fields[0].device)
            mi = torch.zeros(n, dtype=torch.float32, device=fields[0].device)
            pc = torch.zeros(n, 5, dtype=torch.float32, device=fields[0].device)
            _sync_producers([st] + list(fields))    # the zero fills above run on torch's stream, the reduction on the library's
            self.check(L.arc_rad_domain_stats(C.byref(d), ms, n, ptrs, C.c_void_p(int(st.data_ptr()))))
            if morans:
                self.check(L.arc_rad_morans_i(C.byref(d), ms, n, ptrs, C.c_void_p(int(mi.data_ptr()))))
            if percentiles:
                self.check(L.arc_rad_percentiles(C.byref(d), ms, n, ptrs, 5, abi.fptr(perc), C.c_void_p(int(pc.data_ptr()))))
            st, mi, pc = st.cpu().numpy(), mi.cpu().numpy(), pc.cpu().numpy()
        else:
            st, mi, pc = np.zeros((n, 5)), np.zeros(n, np.float32), np.zeros((n, 5), np.float32)
            self.check(L.arc_rad_domain_stats(C.byref(d), ms, n, ptrs, C.c_void_p(st.ctypes.data)))
            if morans:
                self.check(L.arc_rad_morans_i(C.byref(d), ms, n, ptrs, C.c_void_p(mi.ctypes.data)))
            if percentiles:
                self.check(L.arc_rad_percentiles(C.byref(d), ms, n, ptrs, 5, abi.fptr(perc), C.c_void_p(pc.ctypes.data)))
        return decomposition.stats_from_sums(st, names=names, morans_i=mi if morans else None, percentiles=pc if percentiles else None)

    def cal_cldfra1(self, dims, CLDFRA, QV, QC, QI, QS, t_phy, p_phy, F_QV=True, F_QC=True, F_QI=True, F_QS=True, F_ICE_PHY=None,
                    mp_physics=0, cldfra1_flag=None):
        """cal_cldfra1 of module_radiation_driver.F:2886-3122 (icloud = 1) with the Fortran dummy names; F_Qx = None means the
        OPTIONAL argument is not PRESENT.  CLDFRA (and cldfra1_flag, int32) are written for the tile levels kts..kte."""
        d = abi.make_dims(dims) if isinstance(dims, dict) else dims
        arrays = [CLDFRA, QV, QC, QI, QS, t_phy, p_phy, F_ICE_PHY, cldfra1_flag]
        dev = [_is_device(v) for v in arrays if v is not None]
        if any(dev) and not all(dev):
            raise ValueError("mix of host and device arrays")
        ms = abi.ARC_MEM_DEVICE if dev[0] else abi.ARC_MEM_HOST
        flag = lambda f: -1 if f is None else int(bool(f))
        pname = "arc_rad_cal_cldfra1" if self.prefix == "arc_rad_" else "arc_oracle_cal_cldfra1"
        fn = getattr(self.lib, pname)
        fn.restype = C.c_int
        iptr = None
        if cldfra1_flag is not None:
            iptr = C.cast(int(cldfra1_flag.data_ptr()) if hasattr(cldfra1_flag, "data_ptr") else cldfra1_flag.ctypes.data, C.c_void_p)
        common = [_ptr(QV), _ptr(QC), _ptr(QI), _ptr(QS), flag(F_QV), flag(F_QC), flag(F_QI), flag(F_QS), _ptr(t_phy), _ptr(p_phy), _ptr(F_ICE_PHY),
                  int(mp_physics), _ptr(CLDFRA), iptr]
        if self.prefix == "arc_rad_":
            fn.argtypes = [C.POINTER(abi.ArcDims), C.c_int] + [abi.c_fp] * 4 + [C.c_int] * 4 + [abi.c_fp] * 3 + [C.c_int, abi.c_fp, C.c_void_p]
            _sync_producers(arrays)
            self.check(fn(C.byref(d), ms, *common))
        else:
            fn.argtypes = [C.POINTER(abi.ArcDims)] + [abi.c_fp] * 4 + [C.c_int] * 4 + [abi.c_fp] * 3 + [C.c_int, abi.c_fp, C.c_void_p]
            rc = fn(C.byref(d), *common)
            if rc:
                raise RadiationError(rc, "oracle cal_cldfra1")

    def _driver_call(self, name, dims, arrays, args_lib, args_orc, types_lib, types_orc):
        """Common plumbing of the radiation_driver pre-processing entry points: one memory space for all arrays, the product
        library takes (dims, memspace, ...), the oracle (dims, ...)."""
        d = abi.make_dims(dims) if isinstance(dims, dict) else dims
        dev = [_is_device(v) for v in arrays if v is not None]
        if any(dev) and not all(dev):
            raise ValueError("mix of host and device arrays")
        ms = abi.ARC_MEM_DEVICE if dev[0] else abi.ARC_MEM_HOST
        fn = getattr(self.lib, ("arc_rad_" if self.prefix == "arc_rad_" else "arc_oracle_") + name)
        fn.restype = C.c_int
        if self.prefix == "arc_rad_":
            fn.argtypes = [C.POINTER(abi.ArcDims), C.c_int] + types_lib
            _sync_producers(arrays)
            self.check(fn(C.byref(d), ms, *args_lib))
        else:
            fn.argtypes = [C.POINTER(abi.ArcDims)] + types_orc
            rc = fn(C.byref(d), *args_orc)
            if rc:
                raise RadiationError(rc, "oracle " + name)

    def cal_cldfra2(self, dims, CLDFRA, QC, QI, F_QC=True, F_QI=True):
        """cal_cldfra2 of module_radiation_driver.F:2801-2874 (icloud = 2): CLDFRA = 1 where QC + QI > 1e-6, tile levels kts..kte."""
        a = [_ptr(QC), _ptr(QI), int(bool(F_QC)), int(bool(F_QI)), _ptr(CLDFRA)]
        t = [abi.c_fp, abi.c_fp, C.c_int, C.c_int, abi.c_fp]
        self._driver_call("cal_cldfra2", dims, [CLDFRA, QC, QI], a, a, t, t)

    def cal_cldfra3(self, dims, CLDFRA, qv, qc, qi, qs, p, t, rho, XLAND, gridkm):
        """cal_cldfra3 of module_radiation_driver.F:3140-3274 (icloud = 3): CLDFRA out, qc / qi INOUT."""
        a = [_ptr(CLDFRA), _ptr(qv), _ptr(qc), _ptr(qi), _ptr(qs), _ptr(p), _ptr(t), _ptr(rho), _ptr(XLAND), float(gridkm)]
        t_ = [abi.c_fp] * 9 + [C.c_float]
        self._driver_call("cal_cldfra3", dims, [CLDFRA, qv, qc, qi, qs, p, t, rho, XLAND], a, a, t_, t_)

    def ozn_time_int(self, dims, julday, julian, ozmixm, ozmixt, levsiz, num_months):
        """ozn_time_int of module_radiation_driver.F:3993-4098: ozmixm (num_months, jms:jme, levsiz, ims:ime in C order) ->
        ozmixt (jms:jme, levsiz, ims:ime), linear in time between the bracketing mid-month days."""
        a = [int(julday), float(julian), int(levsiz), int(num_months), _ptr(ozmixm), _ptr(ozmixt)]
        t = [C.c_int, C.c_float, C.c_int, C.c_int, abi.c_fp, abi.c_fp]
        self._driver_call("ozn_time_int", dims, [ozmixm, ozmixt], a, a, t, t)

    def ozn_p_int(self, dims, p, pin, levsiz, ozmixt, o3vmr):
        """ozn_p_int of module_radiation_driver.F:4100-4234: ozone on the data levels pin (host float32 array) -> o3vmr at p."""
        pin = np.ascontiguousarray(pin, dtype=np.float32)
        a = [_ptr(p), abi.fptr(pin), int(levsiz), _ptr(ozmixt), _ptr(o3vmr)]
        t = [abi.c_fp, abi.c_fp, C.c_int, abi.c_fp, abi.c_fp]
        self._driver_call("ozn_p_int", dims, [p, ozmixt, o3vmr], a, a, t, t)

    def aer_time_int(self, dims, julday, julian, aerodm, aerodt, levsiz, num_months, no_src):
        """aer_time_int of module_radiation_driver.F:4236-4343: aerodm (no_src, num_months, jms:jme, levsiz, ims:ime in C order) ->
        aerodt (no_src, jms:jme, levsiz, ims:ime)."""
        a = [int(julday), float(julian), int(levsiz), int(num_months), int(no_src), _ptr(aerodm), _ptr(aerodt)]
        t = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, abi.c_fp, abi.c_fp]
        self._driver_call("aer_time_int", dims, [aerodm, aerodt], a, a, t, t)

    def aer_p_int(self, dims, p, pin, levsiz, aerodt, aerod, no_src, pf, totaod):
        """aer_p_int of module_radiation_driver.F:4345-4506: -> AEROD (no_src, jms:jme, kms:kme, ims:ime) and TOTAOD (jms:jme, ims:ime)."""
        pin = np.ascontiguousarray(pin, dtype=np.float32)
        a = [_ptr(p), abi.fptr(pin), int(levsiz), _ptr(aerodt), _ptr(aerod), int(no_src), _ptr(pf), _ptr(totaod)]
        t = [abi.c_fp, abi.c_fp, C.c_int, abi.c_fp, abi.c_fp, C.c_int, abi.c_fp, abi.c_fp]
        self._driver_call("aer_p_int", dims, [p, aerodt, aerod, pf, totaod], a, a, t, t)

    def optical_averaging(self, dims, mode, bins, alt, dz8w, outs, sigmag=None):
        """bins: list (one per size section, or per mode) of dicts {species_name: array, ..., "num": array}; a species name
        starts with its class (so4, no3, cl, nh4, na, oin, oc, bc, water), e.g. "oc_orgaro1j".  outs: dict with
        tauaer300..999, gaer*, waer*, tauaerlw1..16 (+ optional extaerlw1..16)."""
        d = abi.make_dims(dims) if isinstance(dims, dict) else dims
        ai, ao = abi.ArcAerIn(), abi.ArcAerOut()
        arrays = [alt, dz8w] + [v for b in bins for v in b.values()]
        dev = [_is_device(v) for v in arrays]
        if any(dev) and not all(dev):
            raise ValueError("mix of host and device arrays")
        ai.memspace = abi.ARC_MEM_DEVICE if dev[0] else abi.ARC_MEM_HOST
        ai.mode = abi.ARC_AER_SECTIONAL if mode == "sectional" else abi.ARC_AER_MODAL
        ai.nbin = len(bins)
        for b, spec in enumerate(bins):
            names = [k for k in spec if k != "num"]
            ai.nspec[b] = len(names)
            ai.num[b] = _ptr(spec["num"])
            ai.sigmag[b] = float(sigmag[b]) if sigmag is not None else 0.0
            for m, name in enumerate(names):
                ai.cls[b][m] = abi.AER_CLASSES.index(name.split("_")[0])
                ai.mass[b][m] = _ptr(spec[name])
        ai.alt, ai.dz8w = _ptr(alt), _ptr(dz8w)
        for w, wl in enumerate((300, 400, 600, 999)):
            ao.tauaer[w] = _ptr(outs["tauaer%d" % wl]); ao.gaer[w] = _ptr(outs["gaer%d" % wl]); ao.waer[w] = _ptr(outs["waer%d" % wl])
        for w in range(16):
            ao.tauaerlw[w] = _ptr(outs["tauaerlw%d" % (w + 1)])
            ao.extaerlw[w] = _ptr(outs.get("extaerlw%d" % (w + 1)))
        fn = getattr(self.lib, ("arc_aer_optics" if self.prefix == "arc_rad_" else "arc_oracle_aer_optics"))
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(abi.ArcDims), C.POINTER(abi.ArcAerIn), C.POINTER(abi.ArcAerOut)]
        _sync_producers(arrays + list(outs.values()))
        rc = fn(C.byref(d), C.byref(ai), C.byref(ao))
        if rc != 0:
            err = getattr(self.lib, "arc_rad_last_error" if self.prefix == "arc_rad_" else "arc_oracle_aer_last_error")
            err.restype = C.c_char_p
            raise RadiationError(rc, (err() or b"").decode(errors="replace"))


def alloc_aer_outputs(dom, like=None, ext=False):
    nj, nkm, ni = dom["t3d"].shape
    if like is not None:
        import torch
        z = lambda: torch.zeros(nj, nkm, ni, dtype=torch.float32, device=like.device)
    else:
        z = lambda: np.zeros((nj, nkm, ni), np.float32)
    o = {}
    for wl in (300, 400, 600, 999):
        for p in ("tauaer", "gaer", "waer"):
            o["%s%d" % (p, wl)] = z()
    for b in range(16):
        o["tauaerlw%d" % (b + 1)] = z()
        if ext:
            o["extaerlw%d" % (b + 1)] = z()
    return o


_LIB = None


def lib() -> RadLib:
    """The product library (CUDA).  Raises ImportError when it has not been built."""
    global _LIB
    if _LIB is None:
        _LIB = RadLib(LIB_PATH, "arc_rad_")
        L = _LIB.lib
        L.arc_rad_launch_count.restype = C.c_longlong
        L.arc_rad_stream.restype = C.c_void_p
        L.arc_rad_set_overlap.restype = C.c_int
        L.arc_rad_set_overlap.argtypes = [C.c_int]
        L.arc_rad_last_kernel_ms.restype = C.c_float
        L.arc_rad_last_kernel_ms.argtypes = [C.c_char_p]
        L.arc_rad_finalize.restype = None
    return _LIB


def rrtmg_init(p_top, kme, sw_data, lw_data, cp=1004.5, device=-1):
    """rrtmg_swinit + rrtmg_lwinit (SW:11211, LW:12845): read + reduce tables, upload to the GPU."""
    return lib().init(p_top, kme, sw_data, lw_data, cp=cp, device=device)


def RRTMG_SWRAD(dims, **kw):
    return lib().RRTMG_SWRAD(dims, **kw)


def RRTMG_LWRAD(dims, **kw):
    return lib().RRTMG_LWRAD(dims, **kw)


# ------------------------------------------------------------------------------------------------
# convenience used by tests / bench: run one ARC radiation step on a synthetic domain dict

SW_FIELDS_3D = ("t3d", "t8w", "p3d", "p8w", "pi3d", "rho3d", "dz8w", "cldfra3d", "qv3d", "qc3d", "qr3d", "qi3d", "qs3d",
                "qg3d", "re_cloud", "re_ice", "re_snow",
                "tauaer300", "tauaer400", "tauaer600", "tauaer999", "gaer300", "gaer400", "gaer600", "gaer999",
                "waer300", "waer400", "waer600", "waer999")
SW_FIELDS_2D = ("xcoszen", "albedo", "tsk", "xland", "xice", "snow")
LW_FIELDS_3D = ("p8w", "p3d", "pi3d", "dz8w", "t3d", "t8w", "rho3d", "cldfra3d", "qv3d", "qc3d", "qr3d", "qi3d", "qs3d",
                "qg3d", "re_cloud", "re_ice", "re_snow") + tuple("tauaerlw%d" % (b + 1) for b in range(16))
LW_FIELDS_2D = ("emiss", "tsk", "xland", "xice", "snow")


def common_flags(dom, clean_atm_diag=1, aer_ra_feedback=1):
    has_re = "re_cloud" in dom
    return dict(icloud=1, warm_rain=0, is_cammgmp_used=0, has_reqc=int(has_re), has_reqi=int(has_re), has_reqs=int(has_re),
                o3input=0, mp_physics=0, aer_ra_feedback=aer_ra_feedback, progn=0, clean_atm_diag=clean_atm_diag,
                f_qv=1, f_qc=1, f_qr=1, f_qi=1, f_qs=1, f_qg=1, r=float(dom["r"]), g=float(dom["g"]))


def alloc_outputs(dom, which, xp=None, like=None, ext=True):
    """Allocate output arrays (numpy, or torch on like.device) for 'sw' or 'lw'; ext adds the 4th (clean+clear) stream."""
    nj, ni = dom["xcoszen"].shape[-2:] if like is None else like.shape[-2:]
    nkm = dom["t3d"].shape[-2]
    names3 = abi.SW_OUT_3D if which == "sw" else abi.LW_OUT_3D
    names2 = (abi.SW_OUT_2D + (abi.SW_OUT_EXT if ext else ())) if which == "sw" else (abi.LW_OUT_2D + (abi.LW_OUT_EXT if ext else ()))
    namesp = abi.SW_OUT_PROF if which == "sw" else abi.LW_OUT_PROF
    out = {}
    if like is not None:
        import torch
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=like.device)
    else:
        z = lambda *s: np.zeros(s, np.float32)
    for n in names3:
        out[n] = z(nj, nkm, ni)
    for n in names2:
        out[n] = z(nj, ni)
    for n in namesp:
        out[n] = z(nj, nkm + 2, ni)
    return out


def sw_kwargs(dom, outs, **flags):
    kw = {k: dom[k] for k in SW_FIELDS_3D + SW_FIELDS_2D if k in dom}
    kw.update(outs)
    kw.update(dict(solcon=float(dom["solcon"]), aer_opt=0, sf_surface_physics=0))
    kw.update(flags)
    return kw


def lw_kwargs(dom, outs, **flags):
    kw = {k: dom[k] for k in LW_FIELDS_3D + LW_FIELDS_2D if k in dom}
    kw.update(outs)
    kw.update(flags)
    return kw
