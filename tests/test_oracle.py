"""Pins the CPU oracle (test infrastructure) against every known answer the reference offers:
G1 (figures/report_simulation.png table), G3 (figures/report.png table), G2 (sim_data.txt), G4 (solarsystem_nbody.py amplitude ranges), the reference's own pure-numpy
functions run on the oracle's series (tests/golden/ref_functions.npz, tests/golden/make_golden.py),
an independent scipy DOP853 integration, scipy's Lomb-Scargle, numpy lstsq/percentile."""
import json
import os

import numpy as np
import pytest

import oracle
from nbody_ai_b200 import _abi, engine, workloads
from conftest import GOLDEN, golden

README = workloads.c1_params(1)[0]


@pytest.fixture(scope="module")
def readme_run(built):
    times, series, final, ns = oracle.integrate_series(README, 365, 8760)
    return times, series, final


def test_g1_report_table(readme_run):
    """figures/report_simulation.png: every number of the table, to the printed precision."""
    times, series, _ = readme_run
    xs = series[:, 0]
    want = dict(a=[0.045, 0.074, 0.107], P=[3.28, 6.99, 12.05], e=[0.012, 0.003, 0.005],
                ttvmax=[11.3, 13.0, 25.2], n=[111, 52, 31], peak=[(16.4, 0.98), (6.2, 0.85), (3.6, 0.80)])
    for j in range(3):
        col = series[:, 3 + 10 * j: 13 + 10 * j]
        tt = oracle.transit_times(col[:, 0], xs, times)
        ttv, m, b = oracle.ttv_fit(tt)
        assert abs(len(tt) - want['n'][j]) <= 1
        assert round(col[:, 4].mean(), 3) == want['a'][j]
        assert round(m, 2) == want['P'][j]
        assert round(col[:, 5].mean(), 3) == want['e'][j]
        assert abs(col[:, 6].mean() * 180 / np.pi - 90.0) < 0.05
        assert round(oracle.maxavg(ttv) * 1440, 1) == want['ttvmax'][j]
        f, p = oracle.lomb_scargle(np.arange(len(ttv)), ttv, 1. / 180, 0.5)
        k = np.argmax(p)
        assert abs(1 / f[k] - want['peak'][j][0]) < 0.25 and abs(p[k] - want['peak'][j][1]) < 0.01
    rv = np.diff(xs) * 1.496e11 * (8760 / (365 * 86400.))
    f, p = oracle.lomb_scargle(times[1:], rv)
    k = np.argmax(p)
    assert abs(1 / f[k] - 7.0) < 0.05 and abs(p[k] - 0.79) < 0.01
    masses = README[1:, 0] * 1.989e30 / 5.972e24
    assert np.allclose(masses, [88.99, 314.0, 137.3], rtol=2e-4)


def test_g3_report_table(built):
    """figures/report.png — the second report() table the reference ships: 1.12 Msun, 0.25 and 0.1 Mjup at
    3.2888 / 7 d, run with example.py's recipe (Ndays = 30 P1, hourly): every number of the table to its printed
    precision, the transit counts of the O-C panel (30 / 14 epochs) and the peaks of the three periodograms as read
    off the figure (O-C: ~16.5 epochs at 0.93 and ~8 epochs at 0.93; RV: 3.29 d at 0.91, the 7 d bump at 0.09; RV
    trace within +-42 m/s).  (The third table, simulations/report.png, gives its inputs only to table precision —
    a randomize() draw with unknown omega_2 — and pins nothing beyond orders of magnitude.)"""
    from nbody_ai_b200.tools import mearth, mjup, msun
    objs = [{'m': 1.12}, {'m': 0.25 * mjup / msun, 'P': 3.2888, 'inc': 3.14159 / 2, 'e': 0, 'omega': 0},
            {'m': 0.1 * mjup / msun, 'P': 7, 'inc': 3.14159 / 2, 'e': 0, 'omega': 0}]
    P = _abi.objects_to_params(objs)
    ndays, nout = 30 * 3.2888, int(30 * 3.2888 * 24)
    times, series, _, _ = oracle.integrate_series(P, ndays, nout)
    xs = series[:, 0]
    want = dict(mass=[79.45, 31.78], a=[0.045, 0.074], P=[3.29, 7.0], e=[0.001, 0.001], ttvmax=[1.2, 1.3], n=[30, 14],
                peak=[(16.5, 0.93), (8.0, 0.93)])
    for j in range(2):
        col = series[:, 3 + 10 * j: 13 + 10 * j]
        tt = oracle.transit_times(col[:, 0], xs, times)
        ttv, m, b = oracle.ttv_fit(tt)
        assert len(tt) == want['n'][j]
        assert round(P[j + 1, 0] * msun / mearth, 2) == want['mass'][j]
        assert round(col[:, 4].mean(), 3) == want['a'][j] and round(m, 2) == want['P'][j]
        assert round(col[:, 5].mean(), 3) == want['e'][j]
        assert round(oracle.maxavg(ttv) * 1440, 1) == want['ttvmax'][j]
        f, p = oracle.lomb_scargle(np.arange(len(ttv)), ttv, 1. / 180, 0.5)
        k = np.argmax(p)
        assert abs(1 / f[k] - want['peak'][j][0]) < 0.5 and abs(p[k] - want['peak'][j][1]) < 0.01
    rv = np.diff(xs) * 1.496e11 * (nout / (ndays * 86400.))
    f, p = oracle.lomb_scargle(times[1:], rv)
    k = np.argmax(p)
    assert abs(1 / f[k] - 3.29) < 0.01 and abs(p[k] - 0.91) < 0.01
    near7 = (1 / f > 6) & (1 / f < 8)
    assert abs(p[near7].max() - 0.09) < 0.01
    assert 38 < rv.max() < 43 and -43 < rv.min() < -38


def test_g2_sim_data(built):
    """sim_data.txt: 30 planet-1 transit times with sigma = 0.5 min noise (example.py:40-44)."""
    g2 = json.load(open(os.path.join(GOLDEN, "sim_data_g2.json")))
    rows = np.array(g2["rows"])
    objs = eval(g2["header"].lstrip("# "))
    P = _abi.objects_to_params(objs)
    assert P.shape == (3, 7) and P[1, 1] == 3.2888
    ndays = int(np.ceil((rows[-1, 0] + 2) * 3.2888))
    times, series, _, _ = oracle.integrate_series(P, ndays, ndays * 24 + 1)
    tt = oracle.transit_times(series[:, 3], series[:, 0], times)
    res = (rows[:, 1] - tt[rows[:, 0].astype(int)]) * 1440
    assert abs(res.mean()) < 0.3          # 0.5/sqrt(30) = 0.09 expected
    assert 0.3 < res.std() < 0.75         # the injected 0.5 min


def test_reference_functions(readme_run):
    """Oracle restatements == the reference's own transit_times/TTV/maxavg/find_zero (bitwise or
    to round-off), on the series recorded in the fixture."""
    ref = golden("ref_functions.npz")
    times, xs = ref["times"], ref["xs"]
    for j in range(3):
        tt = oracle.transit_times(ref["xp%d" % j], xs, times)
        assert np.array_equal(tt, ref["tt%d" % j])                      # same formula, same order: bit-exact
        ttv, m, b = oracle.ttv_fit(tt)
        assert np.abs(ttv - ref["ttv%d" % j]).max() < 1e-11             # lstsq (SVD) vs centred sums
        assert abs(m - ref["mb%d" % j][0]) < 1e-12 and abs(b - ref["mb%d" % j][1]) < 1e-11
        assert abs(oracle.maxavg(ref["ttv%d" % j]) - float(ref["maxavg%d" % j])) < 1e-16
    # the recorded series is the oracle's: it must reproduce
    assert np.abs(readme_run[1][:, 0] - xs).max() < 1e-15


def test_initial_conditions(built):
    """barycentric x0 of the README system (REBOUND Jacobi add + move_to_com)."""
    st = oracle.initial_state(README)
    want = [-1.12451845018e-4, 0.044837913933536, 0.074296406174462, 0.10653960937946]
    assert np.allclose(st[:, 1], want, rtol=1e-11, atol=0)
    com = (st[:, :1] * st[:, 1:]).sum(0)
    assert np.abs(com).max() < 1e-18


@pytest.mark.parametrize("seed,short", [(3, False), (5, True)])
def test_ias15_vs_dop853(built, seed, short):
    """Independent integrator: scipy DOP853 at rtol 1e-13 on the same initial conditions,
    including short-period systems where the hourly landing steps dominate."""
    from scipy.integrate import solve_ivp
    from nbody_ai_b200.simulation import randomize
    rng = np.random.default_rng(seed)
    G = _abi.NBT_G
    done = 0
    while done < 3:
        obj = randomize(rng)
        if short and obj[1]['P'] > 1.5:
            continue
        if obj[2]['P'] / obj[1]['P'] < 2.0:
            continue
        done += 1
        P = _abi.objects_to_params(obj)
        st = oracle.initial_state(P)
        m = st[:, 0]
        y0 = np.concatenate([st[:, 1:4].ravel(), st[:, 4:7].ravel()])

        def rhs(t, y):
            n = len(m)
            x = y[:3 * n].reshape(n, 3)
            d = x[None, :, :] - x[:, None, :]
            r3 = (d ** 2).sum(-1) ** 1.5
            np.fill_diagonal(r3, np.inf)
            a = (G * m[None, :, None] * d / r3[:, :, None]).sum(1)
            return np.concatenate([y[3 * n:], a.ravel()])

        T = 40.0
        sol = solve_ivp(rhs, (0, T), y0, method='DOP853', rtol=1e-13, atol=1e-18)
        _, _, final, _ = oracle.integrate_series(P, T, int(T * 24) + 1)
        err = np.abs(final[:, :3].ravel() - sol.y[:9, -1]).max()
        assert err < 5e-11, (obj, err)


def test_lomb_scargle_vs_scipy(built):
    from scipy.signal import lombscargle
    rng = np.random.default_rng(0)
    t = np.sort(rng.uniform(0, 100, 200))
    y = np.sin(2 * np.pi * t / 7.3) + 0.3 * rng.normal(size=200) + 4.0
    f, p = oracle.lomb_scargle(t, y, 1. / 50, 0.5)
    try:
        q = lombscargle(t, y - y.mean(), 2 * np.pi * f, floating_mean=True, normalize=True)
    except TypeError:
        pytest.skip("scipy without floating_mean")
    assert np.abs(p - q).max() < 1e-12
    # astropy's autofrequency grid: df = 1/T/spp, Nf = 1 + round-half-even((fmax - fmin)/df)
    T = t.max() - t.min()
    assert len(f) == 1 + int(np.round((0.5 - 1. / 50) / (1.0 / T / 10)))
    assert np.allclose(f, 1. / 50 + (1.0 / T / 10) * np.arange(len(f)), rtol=0, atol=0)


def test_nfreq_round_half_even(built):
    # n = 10 epochs: (0.5 - 1/180) * 10 * 9 = 44.5 in exact arithmetic; follow numpy's rint
    for n in range(3, 400):
        T = n - 1.0
        df = 1.0 / T / 10
        want = 1 + int(np.round((0.5 - 1. / 180) / df))
        assert oracle.lib().nbo_ls_nfreq(T, 1. / 180, 0.5, 10) == want


def test_ttv_fit_vs_lstsq(built):
    rng = np.random.default_rng(1)
    for n in (3, 4, 31, 111, 1000):
        tt = 0.7 + 3.2888 * np.arange(n) + 1e-3 * rng.normal(size=n)
        A = np.vstack([np.ones(n), np.arange(n)]).T
        b, m = np.linalg.lstsq(A, tt, rcond=None)[0]
        ttv, m2, b2 = oracle.ttv_fit(tt)
        assert abs(m - m2) < 1e-12 and abs(b - b2) < 1e-10
        assert np.abs(ttv - (tt - m * np.arange(n) - b)).max() < 1e-10


def test_maxavg_vs_numpy(built):
    rng = np.random.default_rng(2)
    for n in (1, 2, 3, 5, 31, 111, 8759):
        x = rng.normal(size=n)
        want = (np.percentile(np.abs(x), 75) + np.max(np.abs(x))) * 0.5
        assert abs(oracle.maxavg(x) - want) < 1e-15


def test_transit_edge_cases(built):
    """both comparisons are inclusive, so a sample with dx == 0 yields the same
    transit twice; no crossing -> empty; upward crossings are ignored."""
    t = np.arange(6.0)
    xs = np.zeros(6)
    tt = oracle.transit_times(np.array([1., 0., -1., -2., 1., 2.]), xs, t)
    assert np.array_equal(tt, [1.0, 1.0])
    assert len(oracle.transit_times(np.array([1., 2., 3., 4., 5., 6.]), xs, t)) == 0
    assert len(oracle.transit_times(np.array([-1., 1., 2., 3., 4., 5.]), xs, t)) == 0
    tt = oracle.transit_times(np.array([3., 1., -1., -3., 2., -2.]), xs, t)
    assert np.allclose(tt, [1.5, 4.5])


def test_few_transits_fallback(built):
    """simulation.py:205-210: < 3 transits -> P = objects P, t0 = 0, ttv = [0], max = 0."""
    P = README.copy()[None]
    P[0, 3, 1] = 400.0                     # outer planet: at most one transit in 365 d
    plan = engine.Plan(P, 365, 8760, substeps=1, flags=_abi.F_MEANS | _abi.F_LS_OC)
    buf = engine.Buffers(plan)
    oracle.run(plan.cfg, P, buf)
    assert buf.n_tt[2] < 3 and buf.period[2] == 400.0 and buf.t0[2] == 0.0 and buf.ttv_max[2] == 0.0
    assert buf.status[0] & _abi.S_FEW_TRANSITS
    assert buf.ls_oc_nfreq[2] == 0
    assert list(buf.ttv_of(0, 2)) == [0.0]


def test_loglike(built):
    rng = np.random.default_rng(4)
    ttv = 1e-3 * rng.normal(size=40)
    x = np.arange(30.0)
    y = 0.8 + 3.3 * x + ttv[:30] + 1e-4 * rng.normal(size=30)
    yerr = np.full(30, 1e-4)
    want = -0.5 * np.sum(((y - (3.3 * x + 0.8) - ttv[x.astype(int)]) / yerr) ** 2)
    assert abs(oracle.loglike(ttv, 0.8, 3.3, x, y, yerr) - want) < 1e-9 * abs(want)
    # penalty branch: model has fewer transits than the largest observed epoch
    short = ttv[:20]
    pen = -10 * np.sum(((y[:20] - (3.3 * x[:20] + 0.8) - short) / yerr[:20]) ** 2)
    assert abs(oracle.loglike(short, 0.8, 3.3, x, y, yerr) - pen) < 1e-9 * abs(pen)


def test_whfast_preset_and_its_corrector():
    """integrator='whfast' (nbody/simulation.py:39-42: WHFast, corrector 5, safe_mode 0, dt = 1e-3 d) restated: (i) at
    REBOUND's dt it agrees with the IAS15 oracle far inside the parity tolerances; (ii) the symplectic corrector does
    what a correct one must — the error against IAS15 falls by ~1/eps (three orders of magnitude) when it is switched
    on, at several step sizes (a sign or coefficient error makes it worse, not better); (iii) the error of the
    uncorrected map scales as dt^2."""
    par = workloads.c1_params(1)[0]
    cols = [3, 13, 23]
    _, ref, _, _ = oracle.integrate_series(par, 30, 121)
    err = {}
    for dt in (1e-3, 0.01, 0.02):
        for corr in (5, 0):
            _, ser, n = oracle.integrate_series_whfast(par, 30, 121, dt=dt, corrector=corr)
            assert abs(n - (30 / dt + 120)) <= 121          # full steps + one shortened step per output
            err[dt, corr] = np.abs(ser[:, cols] - ref[:, cols]).max()
    assert err[1e-3, 5] < 5e-11 and err[1e-3, 0] < 1e-8
    for dt in (1e-3, 0.01, 0.02):
        assert err[dt, 5] < err[dt, 0] / 200
    assert 3.0 < err[0.02, 0] / err[0.01, 0] < 5.0
    # transit times at the reference's settings: whfast vs IAS15
    t, ser, _ = oracle.integrate_series_whfast(par, 60, 1441)
    _, ias, _, _ = oracle.integrate_series(par, 60, 1441)
    for j in range(3):
        a = oracle.transit_times(ser[:, 3 + 10 * j], ser[:, 0], t)
        b = oracle.transit_times(ias[:, 3 + 10 * j], ias[:, 0], t)
        assert len(a) == len(b) and np.abs(a - b).max() * 86400 < 1e-3
