"""Drop-in for the forward model and likelihood of nbody/nested.py (get_ttv :44-47, the
model_sim/myloglike closures of nlfit :49-95).  The sampler itself (MultiNest, a sequential
Fortran library) is out of scope; what it calls per sample is what the GPU accelerates, and
the batched forms below evaluate B parameter vectors in one call (README.md:94: a retrieval
needs "~5000 N-body simulations or more").
"""
import copy
import ctypes as C

import numpy as np

from . import _abi, engine
from .simulation import analyze, generate, integrate, systems_to_params


def get_ttv(objects, ndays=60, ttvfast=True):
    """nested.py:44-47: integrate at 30-minute cadence and analyze."""
    sim = generate(objects)
    sim_data = integrate(sim, objects, ndays, ndays * 24 * 2)  # dt = 30 minutes
    return analyze(sim_data, ttvfast=ttvfast)


def nlfit_ndays(xx, objects):
    """Integration length used by nlfit (nested.py:60): 1.5 x the span of the observed epochs."""
    return int(np.round(1.5 * (max(xx) + 1) * objects[1]['P']))


def apply_params(objects, bounds, params):
    """model_sim's parameter mapping (nested.py:65-70): params[1:] fill objects[i][k] in the
    order of the bounds dicts.  Returns a new objects list (the reference mutates in place)."""
    obj = copy.deepcopy(objects)
    c = 1
    for i in range(1, len(bounds)):
        for k in bounds[i].keys():
            obj[i][k] = params[c]
            c += 1
    return obj


class ForwardBatch:
    """B forward models (planet-1 transit products only) resident after one nbt_run call."""

    def __init__(self, buffers):
        self.buffers = buffers
        self.plan = buffers.plan

    def __len__(self):
        return self.plan.B

    def __getitem__(self, b):
        """(epochs, ttv, tt) like analyze(ttvfast=True) (simulation.py:181-184)."""
        tt = self.buffers.tt_of(b, 0).copy()
        n = len(tt)
        off = int(self.plan.tt_offsets[b * self.plan.Np])
        ttv = self.buffers.ttv[off:off + n].copy() if n >= 3 else np.zeros(n)
        return np.arange(n), ttv, tt


def get_ttv_batch(systems, ndays=60, device=0, **plan_kw):
    """get_ttv for B systems at once: 30-minute cadence, planet-1 products (ttvfast)."""
    params = systems_to_params(systems)
    flags = _abi.F_PLANET1_ONLY
    plan = engine.Plan(params, ndays, int(ndays * 24 * 2), flags=flags, device=device, **plan_kw)
    return ForwardBatch(engine.run(plan))


def loglike_batch(fwd, c0, c1, xx, yy, yerr, loss="linear"):
    """myloglike (nested.py:85-95) for every forward model of `fwd`:
    -0.5 * sum(loss(((yy - (c1*xx + c0) - ttv[xx]) / yerr)^2)), loss = 'linear' | 'soft_l1' (nested.py:79-82), with
    the reference's penalty branch (-10 * sum over the transits that exist) when an observed epoch is missing
    from the model."""
    lib = _abi.load()
    buf, plan = fwd.buffers, fwd.plan
    B = plan.B
    c0 = np.ascontiguousarray(np.broadcast_to(np.asarray(c0, dtype=np.float64), (B,)))
    c1 = np.ascontiguousarray(np.broadcast_to(np.asarray(c1, dtype=np.float64), (B,)))
    xx, yy, yerr = (np.ascontiguousarray(a, dtype=np.float64) for a in (xx, yy, yerr))
    out = np.zeros(B)
    _abi.check(lib.nbt_chi2(_abi.ptr(buf.ttv), _abi.ptr(plan.tt_offsets, C.c_int64), _abi.ptr(buf.n_tt, C.c_int32),
                            B, plan.Np, _abi.ptr(c0), _abi.ptr(c1), _abi.ptr(xx), _abi.ptr(yy), _abi.ptr(yerr),
                            len(xx), _abi.LOSSES[loss], _abi.ptr(out), plan.cfg.device, 0, None), lib)
    return out


class NestedLikelihood:
    """Vectorised form of nlfit's myloglike (nested.py:49-95): `loglike_batch(cubes)` evaluates B parameter vectors
    cube = [c0 (tmid), c1 (period), free parameters in bounds-dict order...] with ONE nbt_run — forward model at
    30-minute cadence (get_ttv, :44-47), linear ephemeris, O-C and the chi^2 all on the device; only loglike[B] and
    status[B] come back (NBT_F_LOGLIKE).  objects, bounds, xx, yy, yerr as nlfit takes them."""

    def __init__(self, xx, yy, yerr, objects, bounds, myloss='linear', device=0, **plan_kw):
        self.xx, self.yy = np.asarray(xx, dtype=float), np.asarray(yy, dtype=float)
        yerr = np.array(yerr, dtype=float)
        if (yerr == 0).any():
            yerr[yerr == 0] = np.mean(yerr[yerr != 0])               # nested.py:62-63
        self.yerr, self.objects, self.bounds, self.myloss = yerr, copy.deepcopy(objects), bounds, myloss
        self.ndays = nlfit_ndays(self.xx, objects)                   # nested.py:60
        self.device, self.plan_kw = device, plan_kw
        self.base = _abi.objects_to_params(self.objects)
        cols = {"m": _abi.P_M, "P": _abi.P_P, "e": _abi.P_E, "inc": _abi.P_INC, "Omega": _abi.P_OMEGA_NODE,
                "omega": _abi.P_OMEGA_PERI, "f": _abi.P_F}
        self.slots = [(i, cols[k]) for i in range(1, len(bounds)) for k in bounds[i].keys()]   # nested.py:65-70

    def params_for(self, cubes):
        cubes = np.atleast_2d(np.asarray(cubes, dtype=float))
        par = np.repeat(self.base[None], len(cubes), axis=0)
        for c, (i, col) in enumerate(self.slots):
            par[:, i, col] = cubes[:, 1 + c]
        return par

    def loglike_batch(self, cubes):
        cubes = np.atleast_2d(np.asarray(cubes, dtype=float))
        like = engine.Likelihood("nested", x=self.xx, y=self.yy, yerr=self.yerr, c0=cubes[:, 0], c1=cubes[:, 1], loss=self.myloss)
        plan = engine.Plan(self.params_for(cubes), self.ndays, int(self.ndays * 24 * 2), flags=_abi.F_PLANET1_ONLY,
                           device=self.device, likelihood=like, **self.plan_kw)
        buf = engine.Buffers(plan, skip=("tt", "ttv", "n_tt", "period", "t0", "ttv_max"))
        engine.run(plan, buf)
        self.last_status = buf.status
        return buf.loglike.copy()

    def loglike(self, cube):
        return float(self.loglike_batch(np.asarray(cube)[None])[0])
