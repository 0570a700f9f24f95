"""ctypes mirror of include/arc_rad.h (field order must match the header exactly)."""
from __future__ import annotations

import ctypes as C

import numpy as np

c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int)
c_bp = C.POINTER(C.c_ubyte)

ARC_MEM_HOST, ARC_MEM_DEVICE = 0, 1
ARC_VAR_FULL, ARC_VAR_CLEAR, ARC_VAR_CLEAN, ARC_VAR_CLEANCLEAR = 1, 2, 4, 8


class ArcDims(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("ids", "ide", "jds", "jde", "kds", "kde", "ims", "ime", "jms", "jme", "kms", "kme",
                 "its", "ite", "jts", "jte", "kts", "kte")]


class ArcConfig(C.Structure):
    _fields_ = [("cp", C.c_float), ("p_top", C.c_float), ("kme", C.c_int), ("device", C.c_int),
                ("inline_tables", C.c_char_p)]


SW_IN_SCALARS_F = ("radt", "degrad", "declin", "solcon", "xtime", "gmt", "r", "g", "julian")
SW_IN_SCALARS_I = ("julday", "icloud", "warm_rain", "is_cammgmp_used", "has_reqc", "has_reqi", "has_reqs",
                   "o3input", "aer_opt", "no_src", "sf_surface_physics", "mp_physics",
                   "aer_ra_feedback", "progn", "clean_atm_diag",
                   "f_qv", "f_qc", "f_qr", "f_qi", "f_qs", "f_qg", "f_qndrop")
SW_IN_3D = ("t3d", "t8w", "p3d", "p8w", "pi3d", "rho3d", "dz8w", "cldfra3d", "lradius", "iradius",
            "qv3d", "qc3d", "qr3d", "qi3d", "qs3d", "qg3d", "qndrop3d", "o33d", "re_cloud", "re_ice", "re_snow",
            "f_ice_phy", "f_rain_phy",
            "tauaer300", "tauaer400", "tauaer600", "tauaer999", "gaer300", "gaer400", "gaer600", "gaer999",
            "waer300", "waer400", "waer600", "waer999", "aerod", "tauaer3d_sw", "ssaaer3d_sw", "asyaer3d_sw")
SW_IN_2D = ("xcoszen", "albedo", "tsk", "xland", "xice", "snow", "alswvisdir", "alswvisdif", "alswnirdir",
            "alswnirdif", "xlat", "xlong")


class ArcSwIn(C.Structure):
    _fields_ = ([("memspace", C.c_int), ("variant_mask", C.c_int)] +
                [(n, C.c_float) for n in SW_IN_SCALARS_F] + [(n, C.c_int) for n in SW_IN_SCALARS_I] +
                [(n, c_fp) for n in SW_IN_3D] + [(n, c_fp) for n in SW_IN_2D])


SW_OUT_3D = ("rthratensw",)
SW_OUT_2D = ("gsw", "swcf", "coszr", "swupt", "swuptc", "swuptcln", "swdnt", "swdntc", "swdntcln",
             "swupb", "swupbc", "swupbcln", "swdnb", "swdnbc", "swdnbcln", "swvisdir", "swvisdif", "swnirdir",
             "swnirdif", "swddir", "swddni", "swddif")
SW_OUT_PROF = ("swupflx", "swupflxc", "swupflxcln", "swdnflx", "swdnflxc", "swdnflxcln")
SW_OUT_EXT = ("swuptclnc", "swdntclnc", "swupbclnc", "swdnbclnc")


class ArcSwOut(C.Structure):
    _fields_ = [(n, c_fp) for n in SW_OUT_3D + SW_OUT_2D + SW_OUT_PROF + SW_OUT_EXT]


LW_IN_SCALARS_F = ("r", "g", "julian")
LW_IN_SCALARS_I = ("yr", "icloud", "warm_rain", "is_cammgmp_used", "has_reqc", "has_reqi", "has_reqs",
                   "o3input", "mp_physics", "aer_ra_feedback", "progn", "clean_atm_diag",
                   "f_qv", "f_qc", "f_qr", "f_qi", "f_qs", "f_qg", "f_qndrop")
LW_IN_3D = ("p8w", "p3d", "pi3d", "dz8w", "t3d", "t8w", "rho3d", "cldfra3d", "lradius", "iradius",
            "qv3d", "qc3d", "qr3d", "qi3d", "qs3d", "qg3d", "qndrop3d", "o33d", "re_cloud", "re_ice", "re_snow",
            "f_ice_phy", "f_rain_phy")
LW_IN_2D = ("emiss", "tsk", "xland", "xice", "snow")


class ArcLwIn(C.Structure):
    _fields_ = ([("memspace", C.c_int), ("variant_mask", C.c_int)] +
                [(n, C.c_float) for n in LW_IN_SCALARS_F] + [(n, C.c_int) for n in LW_IN_SCALARS_I] +
                [(n, c_fp) for n in LW_IN_3D] + [("tauaerlw", c_fp * 16)] + [(n, c_fp) for n in LW_IN_2D])


LW_OUT_3D = ("rthratenlw",)
LW_OUT_2D = ("glw", "olr", "lwcf", "lwupt", "lwuptc", "lwuptcln", "lwdnt", "lwdntc", "lwdntcln",
             "lwupb", "lwupbc", "lwupbcln", "lwdnb", "lwdnbc", "lwdnbcln")
LW_OUT_PROF = ("lwupflx", "lwupflxc", "lwupflxcln", "lwdnflx", "lwdnflxc", "lwdnflxcln")
LW_OUT_EXT = ("lwuptclnc", "lwdntclnc", "lwupbclnc", "lwdnbclnc")


class ArcLwOut(C.Structure):
    _fields_ = [(n, c_fp) for n in LW_OUT_3D + LW_OUT_2D + LW_OUT_PROF + LW_OUT_EXT]


DBG_I = ("laytrop", "jp", "jt", "jt1", "indfor", "indself", "indminor")
DBG_F1 = ("fac00", "fac01", "fac10", "fac11")
DBG_F2 = ("taug", "taur", "sfluxzen", "taucmc", "hr", "sw_cond")


class ArcDebug(C.Structure):
    _fields_ = ([(n, c_ip) for n in DBG_I] + [(n, c_fp) for n in DBG_F1] + [("cldmask", c_bp)] +
                [(n, c_fp) for n in DBG_F2])


ARC_AER_MAXBIN, ARC_AER_MAXSPEC = 8, 24
ARC_AER_SECTIONAL, ARC_AER_MODAL = 1, 2
AER_CLASSES = ("so4", "no3", "cl", "nh4", "na", "oin", "oc", "bc", "water")


class ArcAerIn(C.Structure):
    _fields_ = [("memspace", C.c_int), ("mode", C.c_int), ("nbin", C.c_int), ("nspec", C.c_int * ARC_AER_MAXBIN),
                ("cls", (C.c_int * ARC_AER_MAXSPEC) * ARC_AER_MAXBIN), ("mass", (c_fp * ARC_AER_MAXSPEC) * ARC_AER_MAXBIN),
                ("num", c_fp * ARC_AER_MAXBIN), ("sigmag", C.c_float * ARC_AER_MAXBIN), ("alt", c_fp), ("dz8w", c_fp)]


class ArcAerOut(C.Structure):
    _fields_ = [("tauaer", c_fp * 4), ("gaer", c_fp * 4), ("waer", c_fp * 4), ("tauaerlw", c_fp * 16), ("extaerlw", c_fp * 16)]


def fptr(a):
    """float* of a numpy float32 C-contiguous array, an int (device address) or None."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.cast(C.c_void_p(int(a)), c_fp)
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], "need contiguous float32"
    return a.ctypes.data_as(c_fp)


def make_dims(d) -> ArcDims:
    return ArcDims(**{k: int(v) for k, v in d.items()})


def alloc_debug(ncol, nlay, ngpt, lw=False):
    """Allocate host tap arrays; returns (ArcDebug, dict of numpy arrays)."""
    a = {
        "laytrop": np.zeros(ncol, np.int32),
        **{k: np.zeros((ncol, nlay), np.int32) for k in ("jp", "jt", "jt1", "indfor", "indself", "indminor")},
        **{k: np.zeros((ncol, nlay), np.float32) for k in DBG_F1},
        "cldmask": np.zeros((ncol, nlay, ngpt), np.uint8),
        "taug": np.zeros((ncol, nlay, ngpt), np.float32),
        "taur": np.zeros((ncol, nlay, ngpt), np.float32),
        "sfluxzen": np.zeros((ncol, ngpt), np.float32),
        "taucmc": np.zeros((ncol, nlay, ngpt), np.float32),
        "hr": np.zeros((ncol, nlay), np.float32),
        "sw_cond": np.ones(ncol, np.float32),
    }
    dbg = ArcDebug()
    for k in DBG_I:
        setattr(dbg, k, a[k].ctypes.data_as(c_ip))
    for k in DBG_F1 + DBG_F2:
        setattr(dbg, k, a[k].ctypes.data_as(c_fp))
    dbg.cldmask = a["cldmask"].ctypes.data_as(c_bp)
    return dbg, a
