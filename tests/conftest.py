import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# parity tolerances of BASELINE.json:north_star
TOL_TT_S = 1.0        # mid-transit times [s]
TOL_OC_MIN = 1e-3     # O-C [min]
TOL_AE_REL = 1e-8     # orbit-averaged a, e (relative)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Compile libnbt.so / the oracle / hostsim if they are missing or stale."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def hostsim(built):
    hs = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"))
    return hs


@pytest.fixture(scope="session")
def nbt(built):
    from nbody_ai_b200 import _abi
    return _abi.load()


@pytest.fixture(scope="session")
def gpu_lib(nbt):
    if nbt.nbt_device_count() <= 0:
        pytest.fail("gpu test collected but libnbt sees no CUDA device")
    return nbt


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def compare_products(got, want, plan, tol_tt=TOL_TT_S, tol_oc=TOL_OC_MIN, tol_ae=TOL_AE_REL, skip=()):
    """Worst-case parity errors between two Buffers of the same plan (got vs want), asserting
    the north-star tolerances.  Returns the worst values."""
    worst = dict(tt=0.0, oc=0.0, a=0.0, e=0.0, inc=0.0)
    for b in range(plan.B):
        if b in skip:
            continue
        for j in range(plan.Np):
            sp = b * plan.Np + j
            assert got.n_tt[sp] == want.n_tt[sp], "transit count differs for system %d planet %d" % (b, j)
            n = got.n_transits(b, j)
            if n:
                worst['tt'] = max(worst['tt'], np.abs(got.tt_of(b, j) - want.tt_of(b, j)).max() * 86400)
            if n >= 3:
                worst['oc'] = max(worst['oc'], np.abs(got.ttv_of(b, j) - want.ttv_of(b, j)).max() * 1440)
            worst['a'] = max(worst['a'], abs(got.mean_a[sp] / want.mean_a[sp] - 1))
            # relative for e, with a floor: for e ~ 1e-6 the element itself is ill-conditioned
            worst['e'] = max(worst['e'], abs(got.mean_e[sp] - want.mean_e[sp]) / max(want.mean_e[sp], 1e-4))
            worst['inc'] = max(worst['inc'], abs(got.mean_inc[sp] - want.mean_inc[sp]))
    assert worst['tt'] < tol_tt, worst
    assert worst['oc'] < tol_oc, worst
    assert worst['a'] < tol_ae, worst
    assert worst['e'] < tol_ae, worst
    return worst
