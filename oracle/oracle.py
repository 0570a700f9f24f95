"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/libarc_oracle.so.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import wrfchem_arc_interactions_b200 as pkg  # noqa: E402
from wrfchem_arc_interactions_b200 import abi  # noqa: E402
from wrfchem_arc_interactions_b200.radiation import RadLib  # noqa: E402

LIB_PATH = os.path.join(_HERE, "libarc_oracle.so")


def build(force=False):
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


_ORC = None


def oracle() -> RadLib:
    global _ORC
    if _ORC is None:
        build()
        _ORC = RadLib(LIB_PATH, "arc_oracle_")
        L = _ORC.lib
        L.arc_oracle_sw_omp.restype = C.c_int
        L.arc_oracle_sw_omp.argtypes = [C.POINTER(abi.ArcDims), C.POINTER(abi.ArcSwIn), C.POINTER(abi.ArcSwOut), C.c_int]
        L.arc_oracle_lw_omp.restype = C.c_int
        L.arc_oracle_lw_omp.argtypes = [C.POINTER(abi.ArcDims), C.POINTER(abi.ArcLwIn), C.POINTER(abi.ArcLwOut), C.c_int]
        L.arc_oracle_table.restype = C.c_int
        L.arc_oracle_table.argtypes = [C.c_int, C.c_int, C.c_char_p, abi.c_fp, C.c_int]
    return _ORC


class _OracleMT(RadLib):
    """Same interface, but every call is spread over `nthreads` host threads by j-rows (static schedule),
    like radiation_driver's OpenMP loop over tiles (module_radiation_driver.F:975-978)."""

    def __init__(self, base: RadLib, nthreads: int):
        self.__dict__.update(base.__dict__)
        L = base.lib
        self.nthreads = int(nthreads)
        self._sw = lambda d, si, so, dbg: L.arc_oracle_sw_omp(d, si, so, self.nthreads)
        self._lw = lambda d, li, lo, dbg: L.arc_oracle_lw_omp(d, li, lo, self.nthreads)


def oracle_mt(nthreads=0) -> RadLib:
    """Oracle bound to the multi-threaded entry points (nthreads <= 0: all hardware threads)."""
    base = oracle()
    n = nthreads if nthreads > 0 else int(base.lib.arc_oracle_max_threads())
    return _OracleMT(base, n)
