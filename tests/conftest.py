import os
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ktab():
    """Synthetic RRTMG_SW_DATA / RRTMG_LW_DATA in the real record layout (the real files are not in the reference)."""
    from wrfchem_arc_interactions_b200 import ktables
    d = tempfile.mkdtemp(prefix="arc_ktab_")
    return ktables.write_files(d)


@pytest.fixture(scope="session")
def orc(ktab):
    import oracle as O
    O.build()
    return O.oracle()


@pytest.fixture(scope="session")
def lib(ktab):
    from wrfchem_arc_interactions_b200 import radiation as R
    return R.lib()


def run_pair(which, rad, dom, debug=None, outs=None, **flag_over):
    """Run RRTMG_SWRAD / RRTMG_LWRAD of `rad` (oracle or CUDA library) on a synthetic domain; returns outputs dict."""
    from wrfchem_arc_interactions_b200 import radiation as R
    flags = R.common_flags(dom)
    flags.update(flag_over)
    outs = outs if outs is not None else R.alloc_outputs(dom, which)
    if which == "sw":
        rad.RRTMG_SWRAD(dom["dims"], debug=debug, **R.sw_kwargs(dom, outs, **flags))
    else:
        rad.RRTMG_LWRAD(dom["dims"], debug=debug, **R.lw_kwargs(dom, outs, **flags))
    return outs


def init(rad, dom, ktab):
    rad.init(dom["p_top"], dom["dims"]["kme"], ktab[0], ktab[1])
    return rad


def interior(dom, a):
    h = dom["halo"]
    if h == 0:
        return a
    return a[h:-h, ..., h:-h] if a.ndim == 3 else a[h:-h, h:-h]
