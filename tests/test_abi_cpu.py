"""The C-ABI library loads without a GPU, exports every symbol include/arc_rad.h declares, and fails loudly (no CPU
fallback) when asked to compute without a CUDA device."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    src = open(os.path.join(ROOT, "include", "arc_rad.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(arc_rad_\w+)\s*\(", src)))


def test_header_declares_the_boundary():
    fns = declared_functions()
    for need in ("arc_rad_init", "arc_rad_sw", "arc_rad_lw", "arc_rad_finalize", "arc_rad_last_error"):
        assert need in fns


def test_library_exports_every_declared_symbol(lib):
    for fn in declared_functions():
        assert hasattr(lib.lib, fn), "libarcrad.so does not export %s" % fn


def test_struct_sizes_match_header(lib):
    """ctypes mirrors must have the layout the C compiler gives the header's structs (checked through a tiny C program)."""
    import subprocess, tempfile
    from wrfchem_arc_interactions_b200 import abi
    d = tempfile.mkdtemp()
    csrc = os.path.join(d, "sz.c")
    open(csrc, "w").write('#include <stdio.h>\n#include "arc_rad.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",sizeof(ArcDims),'
                          'sizeof(ArcConfig),sizeof(ArcSwIn),sizeof(ArcSwOut),sizeof(ArcLwIn),sizeof(ArcLwOut),sizeof(ArcDebug));return 0;}\n')
    exe = os.path.join(d, "sz")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), csrc, "-o", exe])
    sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    mine = [C.sizeof(t) for t in (abi.ArcDims, abi.ArcConfig, abi.ArcSwIn, abi.ArcSwOut, abi.ArcLwIn, abi.ArcLwOut, abi.ArcDebug)]
    assert sizes == mine


def test_calls_before_init_are_refused(lib):
    from wrfchem_arc_interactions_b200 import radiation as R, synth
    if lib.initialised:
        pytest.skip("library already initialised in this process")
    dom = synth.make_domain(4, 2, 8, seed=3)
    outs = R.alloc_outputs(dom, "sw")
    with pytest.raises(R.RadiationError) as e:
        lib.RRTMG_SWRAD(dom["dims"], **R.sw_kwargs(dom, outs, **R.common_flags(dom)))
    assert e.value.code == 1      # ARC_ERR_NOT_INIT


def test_no_cpu_fallback(lib, ktab):
    """Without a CUDA device init must fail with ARC_ERR_CUDA: the product never computes on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from wrfchem_arc_interactions_b200 import radiation as R
    with pytest.raises(R.RadiationError) as e:
        lib.init(5000.0, 41, ktab[0], ktab[1])
    assert e.value.code == 7      # ARC_ERR_CUDA
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "wrfchem-arc-interactions_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "libarc_oracle" not in txt and "import oracle" not in txt and "oracle/" not in txt, f


def test_sweep_groups_partition_the_g_points(lib):
    """Sweep groups (k_sw_sweep / k_lw_sweep): consecutive g-points of one band, every g-point exactly once, sizes balanced and
    <= gmax; with gmax >= the largest band the groups are the bands (the reference's per-band accumulation, LW:3365-3395)."""
    L = lib.lib
    L.arc_rad_test_sweep_groups.restype = C.c_int
    sw = [6, 12, 8, 8, 10, 10, 2, 10, 8, 6, 6, 8, 6, 12]            # ngc, module_ra_rrtmg_sw.F:4816
    lw = [10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2]    # ngc, module_ra_rrtmg_lw.F:8146
    for ng in (sw, lw):
        for gmax in (16, 12, 8, 6, 5):
            band, g0, size = (C.c_int * 32)(), (C.c_int * 32)(), (C.c_int * 32)()
            n = L.arc_rad_test_sweep_groups((C.c_int * len(ng))(*ng), len(ng), gmax, band, g0, size)
            if sum((x + gmax - 1) // gmax for x in ng) > 32:
                assert n == -1          # more groups than the kernels' descriptor table holds: refused
                continue
            assert 0 < n <= 32
            nxt = 0
            per_band = {}
            for q in range(n):
                assert g0[q] == nxt and 1 <= size[q] <= gmax
                nxt += size[q]
                per_band.setdefault(band[q], []).append(size[q])
            assert nxt == sum(ng)
            assert sorted(per_band) == list(range(len(ng)))
            for b, sizes in per_band.items():
                assert sum(sizes) == ng[b] and max(sizes) - min(sizes) <= 1
                assert len(sizes) == (ng[b] + gmax - 1) // gmax


def test_tiled_coefficient_layout_is_a_bijection(lib):
    """coef_index: [layer][32-column tile][field][lane] - every (field, layer, column) maps to its own word, a thread's fields of a
    layer are exactly 32 words apart and the 32 lanes of a tile are contiguous."""
    L = lib.lib
    L.arc_rad_test_coef_index.restype = C.c_longlong
    L.arc_rad_test_coef_index.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_int]
    cap, nf, nl = 512, 15, 7
    seen = set()
    for lay in range(nl):
        for f in range(nf):
            for c in range(cap):
                seen.add(L.arc_rad_test_coef_index(f, lay, c, cap, nf))
    assert len(seen) == cap * nf * nl and min(seen) == 0 and max(seen) == cap * nf * nl - 1
    assert L.arc_rad_test_coef_index(3, 2, 70, cap, nf) - L.arc_rad_test_coef_index(2, 2, 70, cap, nf) == 32
    assert L.arc_rad_test_coef_index(3, 2, 71, cap, nf) - L.arc_rad_test_coef_index(3, 2, 70, cap, nf) == 1
    assert L.arc_rad_test_coef_index(0, 3, 5, cap, nf) - L.arc_rad_test_coef_index(0, 2, 5, cap, nf) == cap * nf
