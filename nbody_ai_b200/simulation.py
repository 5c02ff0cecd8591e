"""Drop-in for nbody/simulation.py of pearsonkyle/Nbody-AI.

Same function names, arguments and return layouts as the reference (generate, integrate,
randomize, transit_times, TTV, analyze), plus the batched forms the reference lacks
(generate_batch / analyze_batch), which are what the GPU is for: thousands to millions of
independent systems per call.  All arithmetic of the path runs in libnbt.so (CUDA, sm_100a)
through the C ABI of include/nbt.h; nothing here falls back to the CPU.

Differences from the reference that a caller can observe:
  * generate() without Ndays/Noutputs returns a light `Simulation` handle (the reference returns
    a rebound.Simulation); integrate(sim, objects, Ndays, Noutputs) accepts it.
  * the trajectory comes from order-(6,4) / order-4 force-gradient Wisdom-Holman compositions whose step is
    planned per system to sit inside the parity tolerances against REBOUND's IAS15
    (DESIGN.md §4), not from IAS15 itself.
  * report() (matplotlib) is out of scope.
"""
import ctypes as C

import numpy as np

from . import _abi, engine
from .tools import (G, au, hill_sphere, lomb_scargle, m2r_star, maxavg, mearth, mjup, msun, rearth, rsun, sa)
from .tools import TTV as _TTV_tools
from .tools import transit_times as _transit_times_tools

SERIES_KEYS = ['x', 'y', 'z', 'P', 'a', 'e', 'inc', 'Omega', 'omega', 'M']  # simulation.py:13-25


def empty_data(N):  # orbit parameters in timeseries — simulation.py:13-25
    return {k: np.zeros(N) for k in SERIES_KEYS}


class Simulation:
    """What generate(objects) returns when no integration is requested (stands in for the
    rebound.Simulation handle of simulation.py:31-52).  Holds a copy of the body parameters, as
    rebound.add() copies them: later mutation of `objects` does not change the handle."""

    def __init__(self, objects, integrator=None):
        if integrator not in (None, 'whfast', 'tes'):
            raise ValueError("unknown integrator %r" % (integrator,))
        self.params = _abi.objects_to_params(objects)
        self.integrator = integrator
        self.N = len(objects)


def _integrator_kw(integrator, Ndays, Noutputs):
    """integrator=None: REBOUND's default (IAS15) is matched within the parity tolerances by the planned
    SBABC3 / C4 compositions (NBT_SCHEME_AUTO).  'whfast' (simulation.py:39-42): NBT_SCHEME_WHFAST — REBOUND's WHFast
    at its default dt = 1e-3 d with the 5th-order symplectic corrector and safe_mode = 0 semantics (corrector applied
    and undone around every sim.integrate(t), last step of an output interval shortened to land on it).  'tes'
    (:34-38) is a high-accuracy scheme: same as None."""
    if integrator == 'whfast':
        return dict(scheme=_abi.SCHEME_WHFAST, substeps=None)
    return dict(scheme=_abi.SCHEME_AUTO, substeps=None)


def generate(objects, Ndays=None, Noutputs=None, integrator=None):
    """simulation.py:27-52.  objects: [{'m': Mstar}, {'m','P','e','inc','omega',...}, ...] in
    Msun / days / radians; bodies are added in Jacobi order and the system moved to its barycentre."""
    sim = Simulation(objects, integrator)
    if Ndays and Noutputs:
        return integrate(sim, objects, Ndays, Noutputs)
    return sim


def integrate(sim, objects, Ndays, Noutputs):
    """simulation.py:131-160: sample the trajectory on np.linspace(0, Ndays, Noutputs)."""
    Noutputs = int(Noutputs)
    plan = engine.Plan(sim.params[None], Ndays, Noutputs, flags=_abi.F_SERIES | _abi.F_MEANS,
                       **_integrator_kw(sim.integrator, Ndays, Noutputs))
    buf = engine.run(plan)
    Np = plan.Np
    ser = buf.series[:, 0, :]
    pdata = []
    for j in range(Np):
        pdata.append({k: np.ascontiguousarray(ser[:, 3 + 10 * j + q]) for q, k in enumerate(SERIES_KEYS)})
    sim_data = ({
        'pdata': pdata,
        'star': {'x': np.ascontiguousarray(ser[:, 0]), 'y': np.ascontiguousarray(ser[:, 1]),
                 'z': np.ascontiguousarray(ser[:, 2])},
        'times': plan.times,
        'objects': objects,
        'dt': Noutputs / (Ndays * 24 * 60 * 60)  # conversion factor to get seconds for RV semi-amp (:157)
    })
    return sim_data


def randomize(rng=None):
    """Random star + 2-planet system from the reference's priors (simulation.py:54-129).
    rng=None uses numpy's global RNG like the reference; pass a np.random.Generator to seed."""
    R = np.random if rng is None else rng
    objects = [
        {'m': R.uniform(0.75, 1.15)},                                   # star mass [msun]
        {'m': R.uniform(0.66 * mearth / msun, 450 * mearth / msun),     # planet 1
         'P': R.uniform(0.8, 20)},
        {'m': R.uniform(0.1, 10),                                       # mass ratio with planet 1
         'omega': R.uniform(0, 2 * np.pi),
         'e': np.abs(R.normal(0, 0.01))},
    ]
    objects[2]['m'] *= objects[1]['m']
    hr1 = hill_sphere(objects, i=1)
    a1 = sa(objects[0]['m'], objects[1]['P'])
    a2, hr2 = 0, 0
    while a1 + hr1 > a2 - hr2:  # period of planet 2 beyond interacting hill spheres
        objects[2]['P'] = objects[1]['P'] * R.uniform(1.25, 10.25)
        a2 = sa(objects[0]['m'], objects[2]['P'])
        hr2 = hill_sphere(objects, i=2)
    objects[1]['inc'] = np.pi / 2
    objects[2]['inc'] = np.pi / 2
    return objects


def transit_times(xp, xs, times):
    """simulation.py:163-170."""
    return _transit_times_tools(xp, xs, times)


def TTV(epochs, tt):
    """simulation.py:172-177: [ttv, m, b] of the linear ephemeris tt ~ b + m*epoch."""
    return _TTV_tools(epochs, tt)


def _series_from_sim_data(m):
    Np = len(m['pdata'])
    NO = len(m['times'])
    ser = np.zeros((NO, 1, 3 + 10 * Np))
    for q, k in enumerate(('x', 'y', 'z')):
        if k in m['star']:
            ser[:, 0, q] = m['star'][k]
    for j in range(Np):
        for q, k in enumerate(SERIES_KEYS):
            if k in m['pdata'][j]:
                ser[:, 0, 3 + 10 * j + q] = m['pdata'][j][k]
    return ser


def analyze(m, ttvfast=False):
    """simulation.py:179-231.  ttvfast=True returns (epochs, ttv, tt) of planet 1 only;
    otherwise the ttv_data dict consumed by report()."""
    if ttvfast:
        tt = transit_times(m['pdata'][0]['x'], m['star']['x'], m['times'])
        ttv, per, t0 = TTV(np.arange(len(tt)), tt)
        return np.arange(len(ttv)), ttv, tt

    lib = _abi.load()
    objects = m['objects']
    times = np.ascontiguousarray(m['times'], dtype=np.float64)
    NO = len(times)
    params = _abi.objects_to_params(objects)[None]
    flags = _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_RV | _abi.F_ORBIT | _abi.F_LS_OC
    plan = engine.Plan(params, float(times[-1] - times[0]), NO, substeps=1, flags=flags)
    # a recorded series need not match `objects` (edited periods, unstable orbits): it can hold at most one transit
    # per sample interval whatever P says, so the capacity is NO - 1 per planet
    Np = plan.Np
    plan.tt_offsets = np.arange(Np + 1, dtype=np.int64) * (NO - 1)
    plan.cfg.tt_cap_max = NO - 1
    lib.nbt_plan_ls(C.byref(plan.cfg), _abi.ptr(plan.tt_offsets, C.c_int64), _abi.ptr(plan.ls_offsets, C.c_int64))
    buf = engine.Buffers(plan)
    series = _series_from_sim_data(m)
    inp, out = buf.as_structs()
    _abi.check(lib.nbt_analyze(C.byref(plan.cfg), _abi.ptr(series), _abi.ptr(times), float(m['dt']),
                               C.byref(inp), C.byref(out), None), lib)
    if buf.status[0] & _abi.S_TT_OVERFLOW:
        raise RuntimeError("analyze(): transit buffer overflow (more than one crossing per sample interval?)")
    RV = buf.rv[:, 0].copy()
    freq, power = lomb_scargle(times[1:], RV)
    data = {
        'times': m['times'],
        'RV': {'freq': freq, 'power': power, 'signal': RV, 'max': float(buf.rv_max[0])},
        'mstar': objects[0]['m'],
        'planets': [],
        'objects': objects,
    }
    for j in range(1, len(objects)):
        data['planets'].append(_planet_dict(buf, 0, j - 1, objects[j]['m']))
    return data


def _planet_dict(buf, b, j, mass):
    """ttv_data['planets'][j] of simulation.py:199-229 from the engine's buffers."""
    p = buf.plan
    sp = b * p.Np + j
    n = buf.n_transits(b, j)
    tt = buf.tt_of(b, j).copy()
    pdata = {}
    if n >= 3:
        ttv = buf.ttv_of(b, j).copy()
    else:
        ttv = [0]  # simulation.py:207-210
    if buf.ls_oc_power is not None:
        if n >= 3:
            freq, power = buf.ls_oc_of(b, j)
            freq, power = freq.copy(), power.copy()
        else:  # LombScargle(arange(1), [0]): zero baseline -> one undefined sample
            freq, power = np.array([np.nan]), np.array([np.nan])
        pdata['freq'], pdata['power'] = freq, power
    pdata['e'] = float(buf.mean_e[sp]); pdata['inc'] = float(buf.mean_inc[sp])
    pdata['a'] = float(buf.mean_a[sp]); pdata['omega'] = float(buf.mean_omega[sp])
    pdata['mass'] = mass
    pdata['P'] = float(buf.period[sp])
    pdata['t0'] = float(buf.t0[sp])
    pdata['tt'] = tt
    pdata['ttv'] = ttv
    pdata['max'] = float(buf.ttv_max[sp])
    if buf.orbit_xz is not None:
        pdata['x'] = buf.orbit_xz[:, b, j, 0].copy()  # x[::4]
        pdata['y'] = buf.orbit_xz[:, b, j, 1].copy()  # z[::4]  (simulation.py:227-228)
    return pdata


# ---------------------------------------------------------------------------------------------
# batched forms (no reference analogue: they replace the per-simulation Python loops of
# simulations/generate_simulations.py:41-110 and simulations/simulation_grid.py:136-153)
# ---------------------------------------------------------------------------------------------
def systems_to_params(systems):
    """[B, n_bodies, 7] parameter block from either an ndarray or a list of `objects` lists."""
    if isinstance(systems, np.ndarray):
        return engine._as_params(systems)
    return np.stack([_abi.objects_to_params(obj) for obj in systems])


class BatchResult:
    """Results of one batched generate+analyze call; `.ttv_data(b)` gives the reference's dict."""

    def __init__(self, buffers, systems=None):
        self.buffers = buffers
        self.plan = buffers.plan
        self.systems = systems

    def __len__(self):
        return self.plan.B

    def ttv_data(self, b):
        buf, p = self.buffers, self.plan
        objects = self.systems[b] if (self.systems is not None and not isinstance(self.systems, np.ndarray)) else \
            [{'m': p.params[b, 0, 0]}] + [dict(m=p.params[b, j, 0], P=p.params[b, j, 1], e=p.params[b, j, 2],
                                               inc=p.params[b, j, 3], Omega=p.params[b, j, 4], omega=p.params[b, j, 5],
                                               f=p.params[b, j, 6]) for j in range(1, p.N)]
        data = {'times': p.times, 'mstar': objects[0]['m'], 'planets': [], 'objects': objects}
        if buf.rv is not None:
            data['RV'] = {'signal': buf.rv[:, b].copy(), 'max': float(buf.rv_max[b])}
            if buf.ls_rv_power is not None:
                data['RV']['freq'], data['RV']['power'] = buf.ls_rv_of(b)
        for j in range(p.Np):
            data['planets'].append(_planet_dict(buf, b, j, objects[j + 1]['m']))
        return data

    def ttvfast(self, b):
        """(epochs, ttv, tt) of planet 1 — analyze(ttvfast=True), simulation.py:181-184."""
        tt = self.buffers.tt_of(b, 0).copy()
        ttv = self.buffers.ttv_of(b, 0).copy() if len(tt) >= 3 else np.zeros(len(tt))
        return np.arange(len(ttv)), ttv, tt


def generate_batch(systems, Ndays, Noutputs, integrator=None, full=True, rv=False, orbit=False,
                   rv_periodogram=False, planet1_only=False, device=0, **plan_kw):
    """generate()+analyze() for B independent systems in one GPU call.

    full=True adds the O-C periodogram and the mean of omega (everything analyze() returns per
    planet); rv / rv_periodogram / orbit add the RV series, its periodogram and x[::4], z[::4]."""
    params = systems_to_params(systems)
    flags = _abi.F_MEANS
    if full:
        flags |= _abi.F_MEAN_OMEGA | _abi.F_LS_OC
    if rv or rv_periodogram:
        flags |= _abi.F_RV
    if rv_periodogram:
        flags |= _abi.F_LS_RV
    if orbit:
        flags |= _abi.F_ORBIT
    if planet1_only:
        flags |= _abi.F_PLANET1_ONLY
    kw = _integrator_kw(integrator, Ndays, Noutputs)
    kw.update(plan_kw)
    plan = engine.Plan(params, Ndays, int(Noutputs), flags=flags, device=device, **kw)
    return BatchResult(engine.run(plan), systems)


def analyze_batch(result):
    """List of ttv_data dicts (analyze() layout) for every system of a BatchResult."""
    return [result.ttv_data(b) for b in range(len(result))]


def report(data, savefile=None):
    """simulation.py:235-334 is a matplotlib figure: plotting is out of scope here, but `data` keeps every key it
    reads (:249-299), so the reference's own report() renders the dict returned by analyze() unchanged.  If the
    reference package is importable (nbody on sys.path, with matplotlib) the call is forwarded to it."""
    try:
        from nbody.simulation import report as reference_report
    except Exception as exc:
        raise NotImplementedError("report() is plotting (out of scope) and the reference's nbody.simulation.report is not "
                                  "importable here (%s); pass this ttv_data dict to it — the layout is identical" % exc)
    return reference_report(data, savefile=savefile)
