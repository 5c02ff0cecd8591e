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


def oracle_buffers(plan):
    import oracle
    from nbody_ai_b200 import engine
    buf = engine.Buffers(plan)
    oracle.run(plan.cfg, plan.params, buf)
    return buf


def oracle_is_chaotic(par_b, n_days, n_outputs, base=None, tol_s=1e-2, tol_e=1e-9):
    """The ONLY parity exclusion: the oracle's own answer moves by more than 0.01 s in a transit time or 1e-9 in a mean
    eccentricity when every period changes by 1e-13 (relative, alternating sign) — or the oracle run itself breaks
    (ejection, collision).  Such a system has no defined parity target: REBOUND and this oracle differ by more than
    that in round-off alone."""
    import oracle
    from nbody_ai_b200 import _abi, engine
    P = np.ascontiguousarray(par_b[None])
    plan = engine.Plan(P, n_days, n_outputs, substeps=1, scheme="sbabc3", flags=_abi.F_MEANS)
    a = base
    if a is None:
        a = engine.Buffers(plan); oracle.run(plan.cfg, P, a, n_threads=1)
    if a.status[0] & (_abi.S_NONFINITE | _abi.S_UNBOUND):
        return True
    Q = P.copy()
    Q[0, 1:, 1] *= 1.0 + 1e-13 * (1 - 2 * (np.arange(P.shape[1] - 1) % 2))
    b = engine.Buffers(plan); oracle.run(plan.cfg, Q, b, n_threads=1)
    if not np.array_equal(a.n_tt, b.n_tt):
        return True
    dtt = max([np.abs(a.tt_of(0, j) - b.tt_of(0, j)).max() * 86400 for j in range(plan.Np) if a.n_transits(0, j)] + [0.0])
    de = np.abs(a.mean_e - b.mean_e) / np.maximum(a.mean_e, 1e-4)
    return bool(dtt > tol_s or de.max() > tol_e)


def parity_report(got, want, plan):
    """Per-system worst differences in units of the parity tolerances (tt 1 s, O-C 1e-3 min, mean a / e 1e-8 relative;
    e with AND without the 1e-4 floor).  Returns arrays [B]: worst (floored), e_rel_nofloor, count_mismatch."""
    B, Np = plan.B, plan.Np
    worst, e_nf, mism = np.zeros(B), np.zeros(B), np.zeros(B, dtype=bool)
    for b in range(B):
        w = 0.0
        for j in range(Np):
            sp = b * Np + j
            if got.n_tt[sp] != want.n_tt[sp]:
                mism[b] = True
                continue
            n = got.n_transits(b, j)
            if n:
                w = max(w, np.abs(got.tt_of(b, j) - want.tt_of(b, j)).max() * 86400 / TOL_TT_S)
            if n >= 3:
                w = max(w, np.abs(got.ttv_of(b, j) - want.ttv_of(b, j)).max() * 1440 / TOL_OC_MIN)
            w = max(w, abs(got.mean_a[sp] / want.mean_a[sp] - 1) / TOL_AE_REL)
            de = abs(got.mean_e[sp] - want.mean_e[sp])
            w = max(w, de / max(want.mean_e[sp], 1e-4) / TOL_AE_REL)
            e_nf[b] = max(e_nf[b], de / want.mean_e[sp] if want.mean_e[sp] > 0 else 0.0)
        worst[b] = np.inf if mism[b] else w
    return worst, e_nf, mism


def chaotic_mask(plan, want, rows, tol_s=1e-2, tol_e=1e-9):
    """oracle_is_chaotic for many systems at once: ONE multi-threaded oracle run over the perturbed copies of systems
    `rows` of `plan`, compared with their unperturbed oracle results in `want`.  Returns a bool array [len(rows)]."""
    import oracle
    from nbody_ai_b200 import _abi, engine
    rows = np.asarray(rows, dtype=np.int64)
    if len(rows) == 0:
        return np.zeros(0, dtype=bool)
    Np = plan.Np
    Q = plan.params[rows].copy()
    Q[:, 1:, 1] *= 1.0 + 1e-13 * (1 - 2 * (np.arange(Np) % 2))
    nd, no = engine._grids_of(plan, rows)
    sub = engine.Plan(Q, nd, no, substeps=1, scheme="sbabc3", flags=(plan.cfg.flags & _abi.F_PLANET1_ONLY) | _abi.F_MEANS, sort=False)
    pert = engine.Buffers(sub)
    oracle.run(sub.cfg, Q, pert)
    out = np.zeros(len(rows), dtype=bool)
    for q, b in enumerate(rows):
        if (want.status[b] | pert.status[q]) & (_abi.S_NONFINITE | _abi.S_UNBOUND):
            out[q] = True
            continue
        if not np.array_equal(want.n_tt[b * Np:(b + 1) * Np], pert.n_tt[q * Np:(q + 1) * Np]):
            out[q] = True
            continue
        dtt = max([np.abs(want.tt_of(b, j) - pert.tt_of(q, j)).max() * 86400 for j in range(Np) if want.n_transits(b, j)] + [0.0])
        ea, eb = want.mean_e[b * Np:(b + 1) * Np], pert.mean_e[q * Np:(q + 1) * Np]
        de = np.abs(ea - eb) / np.maximum(ea, 1e-4)
        out[q] = bool(dtt > tol_s or de.max() > tol_e)
    return out


def assert_parity_or_chaotic(got, want, plan, max_excluded_frac=1.0):
    """Every system inside the tolerances, except those the oracle itself declares chaotic (tested only for the systems
    that miss).  Returns (n_excluded, worst of the compared)."""
    worst, _, _ = parity_report(got, want, plan)
    miss = np.flatnonzero(~(worst < 1.0))
    chaotic = chaotic_mask(plan, want, miss)
    for b, c in zip(miss, chaotic):
        assert c, "system %d misses the tolerances (%.3g x) and is not chaotic" % (b, worst[b])
    excluded = len(miss)
    assert excluded <= max_excluded_frac * plan.B, "%d of %d systems excluded as chaotic" % (excluded, plan.B)
    keep = worst < 1.0
    return excluded, float(worst[keep].max()) if keep.any() else 0.0
