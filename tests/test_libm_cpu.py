"""The product's logf / expf / powf (csrc/glibc_math.cuh, host instantiation of the same source the kernels compile) against
the C library the reference's Fortran calls (through the oracle library, which links libm): bit-exact.

logf: EVERY float in [1e-3, 1200] (the layer pressures in hPa that define jp, SW:2854 / LW:3650) plus a strided sweep of
the whole positive normal range; powf / expf: 2e7 random arguments each over the ranges the path uses and far beyond."""
import ctypes as C

import numpy as np

from wrfchem_arc_interactions_b200 import abi, radiation as R


def _fns(orc):
    L, O = R.lib().lib, orc.lib
    L.arc_rad_selftest_libm.restype = C.c_int
    L.arc_rad_selftest_libm.argtypes = [C.c_int, abi.c_fp, abi.c_fp, C.c_int, abi.c_fp, C.c_int]
    O.arc_oracle_libm.restype = C.c_int
    O.arc_oracle_libm.argtypes = [C.c_int, abi.c_fp, abi.c_fp, C.c_int, abi.c_fp]
    return L, O


def _compare(L, O, which, x, y=None, on_device=0):
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32) if y is not None else x
    a, b = np.empty_like(x), np.empty_like(x)
    assert L.arc_rad_selftest_libm(which, abi.fptr(x), abi.fptr(y), x.size, abi.fptr(a), on_device) == 0
    assert O.arc_oracle_libm(which, abi.fptr(x), abi.fptr(y), x.size, abi.fptr(b)) == 0
    return int((a.view(np.uint32) != b.view(np.uint32)).sum())


def test_logf_every_pressure(orc):
    L, O = _fns(orc)
    lo, hi = int(np.float32(1e-3).view(np.uint32)), int(np.float32(1200.0).view(np.uint32))
    step, n, bad = 1 << 24, 0, 0
    for s0 in range(lo, hi + 1, step):
        bits = np.arange(s0, min(s0 + step, hi + 1), dtype=np.uint32)
        bad += _compare(L, O, 0, bits.view(np.float32)); n += bits.size
    assert n > 169_000_000 and bad == 0, "%d of %d logf results differ from the C library" % (bad, n)


def test_logf_whole_range(orc):
    L, O = _fns(orc)
    bits = np.arange(0x00800000, 0x7f800000, 97, dtype=np.uint32)          # positive normal floats
    assert _compare(L, O, 0, bits.view(np.float32)) == 0
    special = np.array([0.0, -1.0, 1e-45, 1e-39, np.inf, 1.0], np.float32)  # library fall-back (host: the same libm)
    assert _compare(L, O, 0, special) == 0


def test_powf(orc):
    L, O = _fns(orc)
    rng = np.random.default_rng(5)
    n = 20_000_000
    x = np.exp(rng.uniform(np.log(1e-6), np.log(1e6), n)).astype(np.float32)
    y = rng.uniform(-4.0, 4.0, n).astype(np.float32)
    assert _compare(L, O, 2, x, y) == 0
    # the Angstrom bases 0.4 / wavemid (SW:11008) and the LW column-rescaling powers (LW:5212, 6017)
    wl = np.array([3.4615, 2.7885, 2.3247, 2.0461, 1.7840, 1.4625, 1.2703, 1.0101, 0.7016, 0.53325, 0.38815, 0.2990, 0.2316, 8.24], np.float32)
    x = np.repeat(np.float32(0.4) / wl, 200_000)
    y = rng.uniform(-3.0, 6.0, x.size).astype(np.float32)
    assert _compare(L, O, 2, x, y) == 0
    x = rng.uniform(1.0, 50.0, 2_000_000).astype(np.float32)
    for e in (0.65, 0.68, 0.77, 0.79, 1.0 / 3.0):
        assert _compare(L, O, 2, x, np.full(x.size, e, np.float32)) == 0
    # results near / beyond the normal range go through the library on both sides
    x = np.array([1e-30, 1e30, 2.0, 0.5, 0.0, 1e-40], np.float32); y = np.array([3.0, 3.0, 200.0, 200.0, 0.333, 2.0], np.float32)
    assert _compare(L, O, 2, x, y) == 0


def test_expf(orc):
    L, O = _fns(orc)
    rng = np.random.default_rng(6)
    x = rng.uniform(-100.0, 100.0, 20_000_000).astype(np.float32)
    assert _compare(L, O, 1, x) == 0
    x = -np.exp(rng.uniform(np.log(1e-8), np.log(90.0), 5_000_000)).astype(np.float32)
    assert _compare(L, O, 1, x) == 0
