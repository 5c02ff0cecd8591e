"""Verified mode (engine.run_verified) and the IAS15 path, run on the host: tests/hostsim compiles the same
nbt_core.cuh / nbt_ias15.cuh functions the CUDA kernels call per thread (test infrastructure; the product has no CPU
path) and stands in for nbt_run here, so that the host-side logic — which systems are re-checked, how results are
spliced — and the arithmetic of the IAS15 kernel are covered without a GPU.  The GPU forms are in test_gpu_parity.py."""
import ctypes as C

import numpy as np
import pytest

import oracle
from nbody_ai_b200 import _abi, engine, workloads
from conftest import compare_products
from test_hostsim import host_run, oracle_run

# systems of workloads.random_systems(12000, 2, seed=1001) that the planner's step does not resolve although the
# oracle is insensitive to a 1e-13 change of the inputs (found by a host sweep of all 12 000): close pairs, Hill
# separation 1.95 - 3.35; 338 and 4241 are scattering events that no fixed step follows
HARD = [338, 365, 534, 1037, 4241, 8625]


@pytest.fixture()
def host_engine(hostsim, monkeypatch):
    monkeypatch.setattr(engine, "run", lambda plan, buffers=None, stream=None: host_run(hostsim, plan))
    return engine


def test_ias15_matches_oracle(hostsim):
    """NBT_SCHEME_IAS15 is the oracle's algorithm in the product's arithmetic: same steps, results to round-off."""
    P = workloads.c1_params(1)
    plan = engine.Plan(P, 60, 1441, scheme="ias15", flags=_abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_RV | _abi.F_ORBIT)
    got, want = host_run(hostsim, plan), oracle_run(plan)
    w = compare_products(got, want, plan)
    assert w['tt'] < 1e-6 and w['oc'] < 1e-9 and w['a'] < 1e-13 and w['e'] < 1e-11
    assert np.abs(got.rv - want.rv).max() < 1e-9 * np.abs(want.rv).max() and np.abs(got.orbit_xz - want.orbit_xz).max() < 1e-13
    assert np.abs(got.mean_omega - want.mean_omega).max() < 1e-9


def test_run_verified_settles_the_hard_systems(host_engine):
    par = workloads.random_systems(12000, 2, seed=1001)
    easy = [0, 1, 2, 3]
    idx = np.array(easy + HARD)
    plan = engine.Plan(par[idx], 365.0, 8760, flags=_abi.F_MEANS)
    want = oracle_run(plan)
    fast = host_engine.run(plan)
    flagged = np.flatnonzero(fast.status & _abi.S_UNCALIBRATED)
    assert set(flagged) >= set(range(len(easy), len(idx)))            # every hard system carries the flag
    with pytest.raises(AssertionError):
        compare_products(fast, want, plan)                             # and the fast mode misses a tolerance on them
    rep = {}
    got = host_engine.run_verified(plan, report=rep)
    compare_products(got, want, plan)
    assert rep["flagged"] == len(flagged) and rep["accepted_doubling"] + rep["ias15"] == rep["flagged"] and rep["ias15"] >= 2
    # fallback=None leaves the unsettled systems flagged instead
    rep2 = {}
    left = host_engine.run_verified(plan, fallback=None, report=rep2)
    assert rep2["unresolved"] == rep["ias15"] and int(((left.status & _abi.S_UNRESOLVED) != 0).sum()) == rep["ias15"]
    # unflagged systems are untouched
    for b in easy:
        if b not in flagged:
            assert np.array_equal(got.tt_of(b, 0), fast.tt_of(b, 0))


def test_splice_roundtrip():
    par = workloads.random_systems(40, 3, seed=4)
    plan = engine.Plan(par, 100.0, 2401, flags=_abi.F_MEANS | _abi.F_LS_OC | _abi.F_RV)
    a = engine.Buffers(plan)
    rng = np.random.default_rng(0)
    for name, _ in engine.OUTPUT_NAMES:
        arr = getattr(a, name)
        if arr is not None:
            arr[...] = rng.integers(0, 1000, arr.shape)
    rows = np.array([3, 17, 18, 39])
    sub = engine._subset(a, rows)
    b = engine.Buffers(plan)
    engine.splice(b, rows, sub)
    for r in rows:
        for j in range(3):
            lo, hi = plan.tt_offsets[r * 3 + j], plan.tt_offsets[r * 3 + j + 1]
            assert np.array_equal(b.tt[lo:hi], a.tt[lo:hi]) and np.array_equal(b.ttv[lo:hi], a.ttv[lo:hi])
            lo, hi = plan.ls_offsets[r * 3 + j], plan.ls_offsets[r * 3 + j + 1]
            assert np.array_equal(b.ls_oc_power[lo:hi], a.ls_oc_power[lo:hi])
        assert np.array_equal(b.period[r * 3:r * 3 + 3], a.period[r * 3:r * 3 + 3]) and b.status[r] == a.status[r]
        assert np.array_equal(b.rv[:, r], a.rv[:, r])
    untouched = np.setdiff1d(np.arange(40), rows)
    assert not b.status[untouched].any() and not b.rv[:, untouched].any()


def test_training_set_generator_on_the_host_build(hostsim, monkeypatch):
    """dataset.generate_training_set (simulations/generate_simulations.py:41-110) with the host build of the device
    integrator standing in for nbt_run: one launch per block on per-system grids, the rejection rule applied to the
    results, the [X, y] layout, resume from an existing set."""
    from nbody_ai_b200 import dataset

    def host_with_amplitude(plan, buffers=None, stream=None):
        buf = host_run(hostsim, plan)
        for b in range(plan.B):                        # k_fit's maxavg(ttv) (tools.py:22), with the oracle's
            n = buf.n_transits(b, 0)
            if n >= 3:
                off = int(plan.tt_offsets[b * plan.Np])
                buf.ttv_max[b * plan.Np] = oracle.maxavg(buf.ttv[off:off + n])
        return buf

    monkeypatch.setattr(engine, "run", host_with_amplitude)
    st = {}
    X, y = dataset.generate_training_set(samples=5, omegas=2, periods=30, rng=np.random.default_rng(3), stats=st)
    assert len(X) == len(y) == 15 and st["launches"] >= 1 and st["simulations"] == 3 * st["candidates"]
    assert all(len(r) == 7 for r in X) and all(len(t) == 30 for t in y)
    for g in range(0, 15, 3):                          # a draw and its omega variants share everything but omega_2
        assert X[g][:5] == X[g + 1][:5] == X[g + 2][:5] and X[g][6] == X[g + 1][6] and len({X[g][5], X[g + 1][5], X[g + 2][5]}) == 3
        P1 = X[g][1]
        assert np.all(np.diff(y[g]) > 0.9 * P1) and np.all(np.diff(y[g]) < 1.1 * P1) and y[g][0] < 1.05 * P1
    # the accepted draws pass the reference's rule in the oracle too
    par = np.zeros((5, 3, 7))
    for i, g in enumerate(range(0, 15, 3)):
        r = X[g]
        par[i, 0, 0], par[i, 1, 1], par[i, 2, 1], par[i, 2, 5], par[i, 2, 2] = r[0], r[1], r[3], r[5], r[6]
        par[i, 1, 0], par[i, 2, 0] = r[2] * dataset.mearth / dataset.msun, r[4] * dataset.mearth / dataset.msun
        par[i, 1:, 3] = np.pi / 2
    nd, no = dataset.grids(par, 30)
    want = oracle_run(engine.Plan(par, nd, no, flags=_abi.F_PLANET1_ONLY))
    for i in range(5):
        assert want.n_tt[2 * i] >= 30 and want.ttv_max[2 * i] * 1440 <= 180
        assert np.abs(want.tt_of(i, 0)[:30] - y[3 * i]).max() * 86400 < 1.0
    # resume: an existing set is continued, not redrawn
    X2, y2 = dataset.generate_training_set(samples=7, omegas=2, periods=30, rng=np.random.default_rng(4), X=list(X), y=list(y))
    assert len(X2) == 21 and X2[:15] == X
