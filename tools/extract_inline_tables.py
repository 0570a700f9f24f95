#!/usr/bin/env python
"""Extract the inline *numeric data tables* of the reference's RRTMG SW/LW sources
into one small binary container, ``<pkg>/data/rrtmg_inline_tables.bin``.

Only physical data are read (cloud optical-property tables, reference atmosphere,
Planck integrals, g-point reduction maps, ozone / temperature climatology
profiles); no executable code is carried over.  The container travels with the
repo because ``/root/reference`` does not exist on the GPU box.

Sources (``/root/reference/WRF-Chem_code/v3.9.1/phys``):
  module_ra_rrtmg_sw.F : swcldpr 6068-7891, swatmref 2993-3047, swcmbdat 4793-4916,
                         swdatinit 4701-4790, swaerpr 4918-5020, wavemin/wavemax 10117-10122
  module_ra_rrtmg_lw.F : lwcldpr 9858-10496, lwatmref 3812-3973, lwavplank 3975-4678,
                         lwcmbdat 8124-8204, lwdatinit 8012-8122, rtrnmc a0/a1/a2 2958-2969,
                         retab 11424-11441, o3data 12747-12770, PPROF/TPROF 11794-11815

Container layout (little endian):
  8s   magic  b"ARCTBL1\\0"
  u32  n_entries
  n_entries x { 32s name, u32 ndim, 4 x u32 dims, 4 x i32 lower bounds, u64 byte offset }
  float32 payloads, Fortran (column-major) element order.
Integer tables are stored as float32 (all values < 2^24).

Usage:  python tools/extract_inline_tables.py [/root/reference]
"""
import os
import re
import struct
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
PHYS = os.path.join(REF, "WRF-Chem_code", "v3.9.1", "phys")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "wrfchem-arc-interactions_b200", "data", "rrtmg_inline_tables.bin")

# name -> (dims, lower bounds); scalars have dims ()
SW_SHAPES = {
    "extliq1": ((58, 14), (1, 16)), "ssaliq1": ((58, 14), (1, 16)), "asyliq1": ((58, 14), (1, 16)),
    "extice2": ((43, 14), (1, 16)), "ssaice2": ((43, 14), (1, 16)), "asyice2": ((43, 14), (1, 16)),
    "extice3": ((46, 14), (1, 16)), "ssaice3": ((46, 14), (1, 16)), "asyice3": ((46, 14), (1, 16)),
    "fdlice3": ((46, 14), (1, 16)),
    "abari": ((5,), (1,)), "bbari": ((5,), (1,)), "cbari": ((5,), (1,)),
    "dbari": ((5,), (1,)), "ebari": ((5,), (1,)), "fbari": ((5,), (1,)),
    "pref": ((59,), (1,)), "preflog": ((59,), (1,)), "tref": ((59,), (1,)),
    "ngc": ((14,), (1,)), "ngs": ((14,), (1,)), "ngm": ((224,), (1,)), "ngn": ((112,), (1,)),
    "ngb": ((112,), (1,)), "wt": ((16,), (1,)),
    "wavenum1": ((14,), (16,)), "wavenum2": ((14,), (16,)), "delwave": ((14,), (16,)),
    "nspa": ((14,), (16,)), "nspb": ((14,), (16,)),
    "wavemin": ((14,), (1,)), "wavemax": ((14,), (1,)),
    # swaerpr (SW:4918-5020): optical-depth ratio, single-scattering albedo and asymmetry of the six ECMWF aerosol types (iaer = 6)
    "rsrtaua": ((14, 6), (1, 1)), "rsrpiza": ((14, 6), (1, 1)), "rsrasya": ((14, 6), (1, 1)),
}
LW_SHAPES = {
    "absliq1": ((58, 16), (1, 1)), "absice0": ((2,), (1,)), "absice1": ((2, 5), (1, 1)),
    "absice2": ((43, 16), (1, 1)), "absice3": ((46, 16), (1, 1)),
    "absliq0": ((), ()), "abscld1": ((), ()),
    "pref": ((59,), (1,)), "preflog": ((59,), (1,)), "tref": ((59,), (1,)),
    "chi_mls": ((7, 59), (1, 1)),
    "totplnk": ((181, 16), (1, 1)), "totplk16": ((181,), (1,)),
    "a0": ((16,), (1,)), "a1": ((16,), (1,)), "a2": ((16,), (1,)),
    "ngc": ((16,), (1,)), "ngs": ((16,), (1,)), "ngm": ((256,), (1,)), "ngn": ((140,), (1,)),
    "ngb": ((140,), (1,)), "wt": ((16,), (1,)),
    "wavenum1": ((16,), (1,)), "wavenum2": ((16,), (1,)), "delwave": ((16,), (1,)),
    "nspa": ((16,), (1,)), "nspb": ((16,), (1,)),
    "retab": ((95,), (1,)),
    "o3sum": ((31,), (1,)), "ppsum": ((31,), (1,)), "o3win": ((31,), (1,)), "ppwin": ((31,), (1,)),
    "pprof": ((60,), (1,)), "tprof": ((60,), (1,)),
}


def logical_lines(path):
    """Yield comment-stripped statements with '&' continuations joined."""
    buf = ""
    with open(path, "r", errors="replace") as fh:
        for raw in fh:
            line = raw.rstrip("\n")
            if line.lstrip().startswith("#"):
                continue
            # strip trailing comment (no '!' inside strings in the data sections we want)
            if "!" in line:
                line = line[: line.index("!")]
            line = line.strip()
            if not line:
                continue
            if line.startswith("&"):
                line = line[1:].strip()
            if line.endswith("&"):
                buf += line[:-1] + " "
                continue
            buf += line
            yield buf
            buf = ""


NUM = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eEdD][+-]?\d+)?$")


def parse_values(body):
    vals = []
    for tok in body.split(","):
        tok = tok.strip()
        if not tok:
            continue
        rep = 1
        if "*" in tok:
            r, tok = tok.split("*")
            rep = int(r)
        tok = re.sub(r"_(rb|im|r8)$", "", tok.strip())
        if not NUM.match(tok):
            raise ValueError("not a number: %r" % tok)
        vals.extend([float(tok.replace("d", "e").replace("D", "e"))] * rep)
    return vals


def parse_slice(s, lo, n):
    s = s.strip()
    if s == ":":
        return list(range(n))
    if ":" in s:
        a, b = s.split(":")
        a = int(a) if a.strip() else lo
        b = int(b) if b.strip() else lo + n - 1
        return list(range(a - lo, b - lo + 1))
    return int(s) - lo


ASSIGN = re.compile(r"^(\w+)\s*(\(([^)]*)\))?\s*=\s*\(/(.*)/\)$")
SCALAR = re.compile(r"^(\w+)\s*=\s*([-+0-9.eE]+)(_rb)?$")
DATA = re.compile(r"^data\s+(\w+)\s*/(.*)/$", re.I)


def extract(path, shapes):
    arrs = {k: np.full(v[0], np.nan, dtype=np.float64) for k, v in shapes.items()}
    for stmt in logical_lines(path):
        m = DATA.match(stmt)
        if m and m.group(1).lower() in shapes:
            name = m.group(1).lower()
            vals = parse_values(m.group(2))
            a = arrs[name]
            assert a.size == len(vals), (name, a.size, len(vals))
            arrs[name] = np.asarray(vals).reshape(a.shape, order="F")
            continue
        m = SCALAR.match(stmt)
        if m and m.group(1).lower() in shapes and shapes[m.group(1).lower()][0] == ():
            arrs[m.group(1).lower()] = np.asarray(float(m.group(2)))
            continue
        m = ASSIGN.match(stmt)
        if not m or m.group(1).lower() not in shapes:
            continue
        name = m.group(1).lower()
        dims, los = shapes[name]
        vals = np.asarray(parse_values(m.group(4)))
        if m.group(3) is None:
            idx = [list(range(n)) for n in dims]
        else:
            parts = m.group(3).split(",")
            assert len(parts) == len(dims), stmt[:80]
            idx = [parse_slice(p, lo, n) for p, lo, n in zip(parts, los, dims)]
        a = arrs[name]
        sel = tuple(i if isinstance(i, int) else np.asarray(i) for i in idx)
        vec_axes = [i for i in idx if not isinstance(i, int)]
        assert len(vec_axes) == 1 or (len(vec_axes) == len(dims) == 1), stmt[:80]
        assert len(vec_axes[0]) == vals.size, (name, stmt[:60], len(vec_axes[0]), vals.size)
        a[sel] = vals
    for k, a in arrs.items():
        assert not np.isnan(a).any(), "table %s incompletely filled" % k
    return arrs


def main():
    sw = extract(os.path.join(PHYS, "module_ra_rrtmg_sw.F"), SW_SHAPES)
    lw = extract(os.path.join(PHYS, "module_ra_rrtmg_lw.F"), LW_SHAPES)
    entries = []
    for pfx, d, shp in (("sw_", sw, SW_SHAPES), ("lw_", lw, LW_SHAPES)):
        for k in sorted(d):
            dims, los = shp[k]
            entries.append((pfx + k, dims, los, np.asarray(d[k], dtype=np.float32)))
    hdr = 8 + 4 + len(entries) * (32 + 4 + 16 + 16 + 8)
    off = hdr
    blob = bytearray()
    blob += b"ARCTBL1\0" + struct.pack("<I", len(entries))
    payload = bytearray()
    for name, dims, los, a in entries:
        d4 = list(dims) + [1] * (4 - len(dims))
        l4 = list(los) + [1] * (4 - len(los))
        blob += struct.pack("<32sI4I4iQ", name.encode(), len(dims), *d4, *l4, off)
        raw = np.asfortranarray(a).tobytes(order="F")
        payload += raw
        off += len(raw)
    assert len(blob) == hdr
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "wb") as fh:
        fh.write(blob + payload)
    print("wrote %s: %d tables, %d bytes" % (os.path.normpath(OUT), len(entries), len(blob) + len(payload)))


if __name__ == "__main__":
    main()
