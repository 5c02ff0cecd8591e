"""Numerics of the device integrator, run on the host: tests/hostsim compiles the *same*
nbt_core.cuh functions the CUDA kernel calls per thread (test infrastructure; the product has
no CPU path) and is compared with the oracle to the north-star tolerances."""
import ctypes as C

import numpy as np
import pytest

import oracle
from nbody_ai_b200 import _abi, engine, workloads
from conftest import compare_products


def host_run(hs, plan):
    buf = engine.Buffers(plan)
    inp, out = buf.as_structs()
    assert hs.hostsim_run(C.byref(plan.cfg), C.byref(inp), C.byref(out)) == 0
    # hostsim covers k_integrate's arithmetic only; the ephemeris fit (k_fit on the GPU) is
    # filled in with the oracle's so that O-C can be compared
    for b in range(plan.B):
        for j in range(plan.Np):
            n = buf.n_transits(b, j)
            if n >= 3:
                off = int(plan.tt_offsets[b * plan.Np + j])
                buf.ttv[off:off + n], buf.period[b * plan.Np + j], buf.t0[b * plan.Np + j] = oracle.ttv_fit(buf.tt[off:off + n])
    return buf


def oracle_run(plan):
    buf = engine.Buffers(plan)
    oracle.run(plan.cfg, plan.params, buf)
    return buf


def test_kepler_drift_two_body(hostsim):
    """Universal-variable drift vs the closed-form ellipse, e up to 0.6, both signs of h,
    steps up to 0.4 P (exercises the safeguarded path too)."""
    hostsim.hostsim_kepler.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double)]
    rng = np.random.default_rng(0)
    mu = _abi.NBT_G * 1.1
    for _ in range(300):
        a, e = rng.uniform(0.02, 0.5), rng.uniform(0, 0.6)
        M0 = rng.uniform(0, 2 * np.pi)
        n = np.sqrt(mu / a ** 3)
        h = rng.uniform(-0.4, 0.4) * 2 * np.pi / n

        def state(M):
            E = M
            for _ in range(60):
                E -= (E - e * np.sin(E) - M) / (1 - e * np.cos(E))
            x, y = a * (np.cos(E) - e), a * np.sqrt(1 - e * e) * np.sin(E)
            r = a * (1 - e * np.cos(E))
            vx, vy = -a * n * np.sin(E) * a / r, a * n * np.sqrt(1 - e * e) * np.cos(E) * a / r
            return np.array([x, y, 0.0, vx, vy, 0.0])

        s = state(M0).copy()
        st = hostsim.hostsim_kepler(mu, h, s.ctypes.data_as(C.POINTER(C.c_double)))
        want = state(M0 + n * h)
        assert st == 0
        assert np.abs(s[:3] - want[:3]).max() < 2e-13 * a
        assert np.abs(s[3:] - want[3:]).max() < 2e-13 * a * n * 3


@pytest.mark.parametrize("scheme,k", [("sbabc3", 1), ("m64", 1), ("m64", 2), ("y4", 4), ("saba4", 8), ("c4", 3)])
def test_readme_parity(hostsim, scheme, k):
    P = workloads.c1_params(1)
    plan = engine.Plan(P, 365, 8760, substeps=k, scheme=scheme, flags=_abi.F_MEANS | _abi.F_MEAN_OMEGA)
    got, want = host_run(hostsim, plan), oracle_run(plan)
    w = compare_products(got, want, plan)
    assert list(got.n_tt) == [111, 52, 31]
    if scheme in ("m64", "sbabc3"):
        assert w['tt'] < 1e-3 and w['oc'] < 1e-6 and w['e'] < 1e-9


def test_planned_substeps_random_systems(hostsim):
    """Planner-chosen k on seeded draws of the reference priors, all tolerances."""
    par = workloads.random_systems(24, 2, seed=11)
    par = par[par[:, 2, 1] / par[:, 1, 1] > 1.8]          # regular systems only
    plan = engine.Plan(par, 200.0, 4801, flags=_abi.F_MEANS)
    got, want = host_run(hostsim, plan), oracle_run(plan)
    compare_products(got, want, plan)
    assert plan.order is None or sorted(plan.order) == list(range(plan.B))
    # default scheme (AUTO): the finely resolved systems took C4 steps, the others SBABC3
    k = plan.substeps
    assert (k < 0).any() and (k > 0).any() and plan.cfg.alt_systems == (k < 0).sum()
    assert np.all(k[plan.order if plan.order is not None else np.arange(plan.B)][-plan.cfg.alt_systems:] < 0)


def test_series_matches_oracle(hostsim):
    """NBT_F_SERIES: the sim_data columns (x,y,z,P,a,e,inc,Omega,omega,M) against the oracle."""
    P = workloads.c1_params(1)
    P[0, 2, 2], P[0, 2, 5] = 0.05, 1.0          # an eccentric planet so omega / M are defined
    plan = engine.Plan(P, 30, 721, substeps=2, flags=_abi.F_SERIES | _abi.F_MEANS)
    got, want = host_run(hostsim, plan), engine.Buffers(plan)
    oracle.run(plan.cfg, P, want)
    d = np.abs(got.series - want.series)
    W = got.series.shape[2]
    for j in range(3):
        c = 3 + 10 * j
        assert d[:, :, c:c + 3].max() < 1e-10                       # x, y, z
        assert (d[:, :, c + 3] / want.series[:, :, c + 3]).max() < 1e-9     # P
        assert (d[:, :, c + 4] / want.series[:, :, c + 4]).max() < 1e-9     # a
        assert d[:, :, c + 5].max() < 1e-9 and d[:, :, c + 6].max() < 1e-9  # e, inc
    c = 3 + 10 * 1
    assert d[:, :, c + 8].max() < 1e-6 and np.abs(np.angle(np.exp(1j * (got.series[:, :, c + 9] - want.series[:, :, c + 9])))).max() < 1e-6
    assert d[:, :, :3].max() < 1e-12                                 # star


def test_refined_mode_is_closer_to_true_root(hostsim):
    """NBT_TT_REFINED is an extra output, not the parity output: it differs from the grid-linear
    time by the interpolation error (up to ~0.4 s for the README system)."""
    P = workloads.c1_params(1)
    a = host_run(hostsim, engine.Plan(P, 100, 2401, substeps=1, flags=_abi.F_MEANS))
    b = host_run(hostsim, engine.Plan(P, 100, 2401, substeps=1, flags=_abi.F_MEANS, tt_mode=_abi.TT_REFINED))
    fine = host_run(hostsim, engine.Plan(P, 100, 100 * 24 * 60 + 1, substeps=1, flags=_abi.F_MEANS))
    assert list(a.n_tt) == list(b.n_tt) == list(fine.n_tt)
    da = np.abs(a.tt_of(0, 0) - fine.tt_of(0, 0)).max() * 86400
    db = np.abs(b.tt_of(0, 0) - fine.tt_of(0, 0)).max() * 86400
    assert 0.05 < da < 1.0 and db < 0.02


def test_scheme_coefficients(built):
    """Order conditions of the BAB compositions (tools/derive_schemes.py): sum a = sum b = 1 and
    the O(eps dt^2), O(eps dt^4) and O(eps^2 dt^2) error terms vanish for M64."""
    a = [-0.0437514219173741137407351, 0.5437514219173741137407351] * 2
    a = [a[0], a[1], a[1], a[0]]
    b = [0.5316386245813511790852625, -0.3086019704406066393237719, 0.5539266917185109204770189,
         -0.3086019704406066393237719, 0.5316386245813511790852625]
    assert abs(sum(a) - 1) < 1e-15 and abs(sum(b) - 1) < 1e-15
    tau = np.concatenate([[0], np.cumsum(a)])
    c = tau - 0.5
    for j in (2, 4):
        assert abs(sum(bi * ci ** j for bi, ci in zip(b, c)) - 1 / ((j + 1) * 2 ** j)) < 1e-15
    e2 = sum(b[i] * b[j] * (tau[j] - tau[i]) for i in range(5) for j in range(i + 1, 5))
    assert abs(e2 - 1 / 6) < 1e-15


def test_gradient_scheme_coefficients(built):
    """Force-gradient compositions: the kicks are a Lobatto quadrature (O(eps dt^2) and, for SBABC3,
    O(eps dt^4) vanish) and the gradient kicks supply what the plain kicks lack of the O(eps^2 dt^2)
    condition: sum_{i<j} b_i b_j (tau_j - tau_i) + sum_i b_i g_i = 1/6."""
    s5 = 5 ** 0.5
    schemes = {
        "c4": ([0.5, 0.5], [1 / 6, 2 / 3, 1 / 6], [0, 1 / 24, 0], (2,)),
        "sbabc3": ([(1 - 1 / s5) / 2, 1 / s5, (1 - 1 / s5) / 2], [1 / 12, 5 / 12, 5 / 12, 1 / 12],
                   [(13 - 5 * s5) / 24, 0, 0, (13 - 5 * s5) / 24], (2, 4)),
    }
    for name, (a, b, g, moments) in schemes.items():
        assert abs(sum(a) - 1) < 1e-15 and abs(sum(b) - 1) < 1e-15
        tau = np.concatenate([[0], np.cumsum(a)])
        c = tau - 0.5
        for j in moments:
            assert abs(sum(bi * ci ** j for bi, ci in zip(b, c)) - 1 / ((j + 1) * 2 ** j)) < 1e-15, name
        n = len(b)
        e2 = sum(b[i] * b[j] * (tau[j] - tau[i]) for i in range(n) for j in range(i + 1, n))
        assert abs(e2 + sum(bi * gi for bi, gi in zip(b, g)) - 1 / 6) < 1e-15, name


def test_inclined_three_dimensional_systems(hostsim):
    """Mutually inclined orbits with arbitrary nodes, arguments of pericentre and starting anomalies:
    exercises the y/z components, the node-dependent element branches and the general acos path of the
    inclination (|cos i| > 0.03)."""
    rng = np.random.default_rng(21)
    B = 10
    par = np.zeros((B, 4, 7))
    par[:, 0, 0] = rng.uniform(0.8, 1.2, B)
    for j, per in enumerate((4.1, 9.7, 23.0)):
        par[:, j + 1, 0] = rng.uniform(3e-6, 6e-4, B)
        par[:, j + 1, 1] = per * rng.uniform(0.95, 1.05, B)
        par[:, j + 1, 2] = rng.uniform(0.0, 0.08, B)
        par[:, j + 1, 3] = np.radians(rng.uniform(70, 110, B))
        par[:, j + 1, 4] = rng.uniform(0, 2 * np.pi, B)
        par[:, j + 1, 5] = rng.uniform(0, 2 * np.pi, B)
        par[:, j + 1, 6] = rng.uniform(0, 2 * np.pi, B)
    plan = engine.Plan(par, 150.0, 3601, flags=_abi.F_MEANS | _abi.F_MEAN_OMEGA)
    got, want = host_run(hostsim, plan), oracle_run(plan)
    w = compare_products(got, want, plan)
    assert w['inc'] < 1e-9
    assert np.abs(got.mean_omega - want.mean_omega).max() < 1e-6
    assert np.abs(np.degrees(got.mean_inc) - 90).max() > 5          # really inclined


def test_branch_free_acos(hostsim):
    """nbt_acos (one code path for both ranges) against libm over [-1, 1], including the ends."""
    x = np.concatenate([np.linspace(-1, 1, 200001), [-1.0, -0.5, 0.5, 1.0, 1 - 1e-15, -1 + 1e-15, 0.0, 1e-300]])
    y = np.zeros_like(x)
    hostsim.hostsim_acos(x.ctypes.data_as(C.POINTER(C.c_double)), y.ctypes.data_as(C.POINTER(C.c_double)), len(x))
    assert np.abs(y - np.arccos(x)).max() < 1e-15


def test_whfast_preset_device_arithmetic_vs_oracle(hostsim):
    """NBT_SCHEME_WHFAST (nbt_whfast.cuh: integrator='whfast', corrector 5, safe_mode 0, dt = 1e-3 d) in the product's
    arithmetic against the oracle's independent restatement of the same algorithm: positions to 1e-12 AU, transit
    times to 1e-5 s, elements to 1e-10 — and against IAS15 inside the parity tolerances."""
    par = np.concatenate([workloads.c1_params(1), workloads.random_systems(3, 3, seed=5)])
    plan = engine.Plan(par, 20.0, 481, scheme="whfast", flags=_abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_SERIES, lanes=0)
    assert list(plan.substeps) == [41] * 4
    got, want = host_run(hostsim, plan), oracle_run(plan)
    assert np.abs(got.series[:, :, :3] - want.series[:, :, :3]).max() < 1e-13
    for j in range(3):
        c = 3 + 10 * j
        assert np.abs(got.series[:, :, c:c + 3] - want.series[:, :, c:c + 3]).max() < 1e-12
        assert np.abs(got.series[:, :, c + 5] - want.series[:, :, c + 5]).max() < 1e-10
    assert np.array_equal(got.n_tt, want.n_tt)
    ias = oracle_run(engine.Plan(par, 20.0, 481, scheme="ias15", flags=_abi.F_MEANS | _abi.F_MEAN_OMEGA, lanes=0))
    for b in range(4):
        for j in range(3):
            if got.n_transits(b, j):
                assert np.abs(got.tt_of(b, j) - want.tt_of(b, j)).max() * 86400 < 1e-5
                assert np.abs(got.tt_of(b, j) - ias.tt_of(b, j)).max() * 86400 < 1e-3
    assert np.all(np.abs(got.mean_e - want.mean_e) / np.maximum(want.mean_e, 1e-4) < 1e-9)
    assert np.all(np.abs(got.mean_e - ias.mean_e) / np.maximum(ias.mean_e, 1e-4) < 1e-8)
