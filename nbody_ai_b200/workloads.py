"""Synthetic inputs of the BASELINE.json configurations, as parameter blocks
[B, n_bodies, 7] = (m, P, e, inc, Omega, omega, f).  Host-side input generation only.

  C1  README example (README.md:23-28)
  C2  random 2-planet systems from the priors of randomize() (nbody/simulation.py:54-129), with the
      omega sweep of simulations/generate_simulations.py:83-107 (1 + `omegas` simulations per draw)
  C3  WASP-18-style nested-sampling forward-model batch (nbody/nested.py:49-95; nl_fit.py:66-79)
  C4  perturber mass x period grid (simulations/simulation_grid.py:88-153)
  C5  random 4-planet systems: randomize() priors extended outward planet by planet
"""
import numpy as np

from . import _abi
from .tools import G, mearth, mjup, msun

HALF_PI = np.pi / 2


def _sa(m, P):
    return (G * m * P ** 2 / (4 * np.pi ** 2)) ** (1. / 3)


def hill_sphere_batch(mstar, m, P):
    """Vectorised nbody/tools.py:38-59: root of a_star + a_planet - a_cent on [0.5, 1.5] x the
    (m/3M)^(1/3) a estimate, by bisection to machine precision."""
    a = _sa(mstar, P)
    h0 = a * (m / (3 * mstar)) ** (1. / 3)
    f = lambda r: G * mstar / (a + r) ** 2 + G * m / r ** 2 - G * mstar / a ** 2
    lo, hi = 0.5 * h0, 1.5 * h0
    flo = f(lo)
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        fm = f(mid)
        same = np.sign(fm) == np.sign(flo)
        lo = np.where(same, mid, lo)
        hi = np.where(same, hi, mid)
    return 0.5 * (lo + hi)


def c1_objects():
    return [
        {'m': 1.12},
        {'m': 0.28 * mjup / msun, 'P': 3.2888, 'inc': 3.14159 / 2, 'e': 0, 'omega': 0},
        {'m': 0.988 * mjup / msun, 'P': 7, 'inc': 3.14159 / 2, 'e': 0, 'omega': 0},
        {'m': 0.432 * mjup / msun, 'P': 12, 'inc': 3.14159 / 2, 'e': 0, 'omega': 0},
    ]


def c1_params(replicas=1):
    return np.repeat(_abi.objects_to_params(c1_objects())[None], replicas, axis=0)


M_PLANET_MAX = 4500 * mearth / msun  # largest planet-2 mass the reference's priors allow (450 Mearth x 10)


def _outer_planet(rng, mstar, m_in, P_in, max_tries=200):
    """Draw (m, P, e, omega) of the next planet outward like planet 2 of randomize(): mass ratio
    U(0.1, 10) (capped at M_PLANET_MAX so that chains of planets stay planetary), P ratio
    U(1.25, 10.25) redrawn until the Hill spheres do not overlap (simulation.py:84-94).  The
    rejection loop is bounded: after max_tries the widest ratio 10.25 is taken."""
    n = len(mstar)
    m = np.minimum(m_in * rng.uniform(0.1, 10, n), M_PLANET_MAX)
    omega = rng.uniform(0, 2 * np.pi, n)
    e = np.abs(rng.normal(0, 0.01, n))
    a1 = _sa(mstar, P_in)
    hr1 = hill_sphere_batch(mstar, m_in, P_in)
    P = P_in * 10.25
    todo = np.ones(n, dtype=bool)
    for _ in range(max_tries):
        if not todo.any():
            break
        idx = np.nonzero(todo)[0]
        Pn = P_in[idx] * rng.uniform(1.25, 10.25, len(idx))
        a2 = _sa(mstar[idx], Pn)
        hr2 = hill_sphere_batch(mstar[idx], m[idx], Pn)
        ok = ~(a1[idx] + hr1[idx] > a2 - hr2)
        P[idx[ok]] = Pn[ok]
        todo[idx[ok]] = False
    return m, P, e, omega


def random_systems(n, n_planets=2, seed=0):
    """n draws of randomize() (n_planets=2), or its outward extension (n_planets>2, config C5)."""
    rng = np.random.default_rng(seed)
    par = np.zeros((n, n_planets + 1, _abi.NBT_NPAR))
    mstar = rng.uniform(0.75, 1.15, n)
    par[:, 0, 0] = mstar
    par[:, 1, 0] = rng.uniform(0.66 * mearth / msun, 450 * mearth / msun, n)
    par[:, 1, 1] = rng.uniform(0.8, 20, n)
    par[:, 1:, 3] = HALF_PI
    for j in range(2, n_planets + 1):
        m, P, e, om = _outer_planet(rng, mstar, par[:, j - 1, 0], par[:, j - 1, 1])
        par[:, j, 0], par[:, j, 1], par[:, j, 2], par[:, j, 5] = m, P, e, om
    return par


def c2_params(n_draws=10000, omegas=6, seed=0):
    """Config C2: n_draws random 2-planet systems, each followed by `omegas` copies with omega_2
    from linspace(-pi, pi, omegas) + N(0, pi/omegas) (generate_simulations.py:83-90)."""
    base = random_systems(n_draws, 2, seed)
    if omegas <= 0:
        return base
    rng = np.random.default_rng(seed + 1)
    out = np.repeat(base, omegas + 1, axis=0).reshape(n_draws, omegas + 1, 3, _abi.NBT_NPAR)
    sweep = np.linspace(-np.pi, np.pi, omegas)[None, :] + rng.normal(0, np.pi / omegas, (n_draws, omegas))
    out[:, 1:, 2, 5] = sweep
    return out.reshape(n_draws * (omegas + 1), 3, _abi.NBT_NPAR)


def c3_params(n=5000, seed=1):
    """Config C3: WASP-18-like star (1.22 Msun, nbody/tools.py:30) + hot Jupiter (P1 = 0.94145 d,
    10.4 Mjup) + a perturber drawn uniformly in bounds shaped like nl_fit.py:66-79."""
    rng = np.random.default_rng(seed)
    par = np.zeros((n, 3, _abi.NBT_NPAR))
    par[:, 0, 0] = 1.22
    par[:, 1, 0] = 10.4 * mjup / msun
    par[:, 1, 1] = 0.94145 + rng.uniform(-1e-4, 1e-4, n)
    par[:, 1:, 3] = HALF_PI
    par[:, 2, 0] = rng.uniform(5, 140, n) * mearth / msun
    par[:, 2, 1] = rng.uniform(1.9, 2.4, n)
    par[:, 2, 2] = rng.uniform(0, 0.1, n)
    par[:, 2, 5] = rng.uniform(0, 2 * np.pi, n)
    return par, dict(ndays=79, noutputs=79 * 48, epochs=56)


# astropy.constants values the reference's grid script overrides tools.py's with (simulation_grid.py:83-85;
# IAU 2015 / CODATA 2018) [3P]
MSUN_ASTROPY, MEARTH_ASTROPY, MJUP_ASTROPY = 1.988409870698051e30, 5.972167867791379e24, 1.8981245973360505e27


def c4_objects():
    """The objects list of simulations/simulation_grid.py:88-113 (planet 2's P is set per cell)."""
    me = MEARTH_ASTROPY / MSUN_ASTROPY
    return [
        {'m': 1},
        {'m': 10 * me, 'P': 1.42, 'inc': np.deg2rad(90), 'e': 0, 'omega': 0.01},
        {'m': 10 * me, 'omega': np.deg2rad(90), 'e': 0, 'inc': np.deg2rad(90)},
    ]


def c4_grids(npts=256):
    """mass_grid, period_grid of simulation_grid.py:115-119: np.meshgrid(mass_range, period_range), i.e. cell [i][j]
    has period_range[i] and mass_range[j]."""
    obj = c4_objects()
    mass_range = np.linspace(0.5, 325, npts) * MEARTH_ASTROPY / MSUN_ASTROPY
    period_range = np.linspace(1.41, 2.41, npts) * obj[1]['P']
    return np.meshgrid(mass_range, period_range)


def c4_params(npts=256, epochs=500):
    """Config C4: the reference's grid (simulation_grid.py:88-153) — 1 Msun star, 10 Mearth inner planet at 1.42 d
    (omega = 0.01), perturber (omega = 90 deg) over mass 0.5..325 Mearth x period ratio 1.41..2.41, astropy constants;
    `epochs` inner periods at 30-minute cadence (generate_sim :23-25).  Cells in row-major order of the meshgrid."""
    obj = c4_objects()
    mass_grid, period_grid = c4_grids(npts)
    obj[2]['P'] = 1.0
    base = _abi.objects_to_params(obj)
    par = np.repeat(base[None], npts * npts, axis=0)
    par[:, 2, _abi.P_M] = mass_grid.ravel()
    par[:, 2, _abi.P_P] = period_grid.ravel()
    ndays = obj[1]['P'] * epochs
    return par, dict(ndays=ndays, noutputs=int(round(obj[1]['P'] * epochs * 48)), epochs=epochs)


def c5_params(n=1000000, seed=2):
    return random_systems(n, 4, seed)
