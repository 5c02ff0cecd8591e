"""Drop-in for the likelihood of ttv_fit.py's `nbody_fitter` (:13-94).  UltraNest itself is out of
scope; it accepts vectorised likelihoods (`ReactiveNestedSampler(..., vectorized=True)`), and that is
the batch boundary: `loglike_batch(thetas)` evaluates B parameter vectors with one nbt_run and one
nbt_ttvfit_loglike instead of B sequential generate() + transit_times() calls."""
import copy
import ctypes as C

import numpy as np

from . import _abi, engine


class NbodyLikelihood:
    """data, prior, bounds exactly as `nbody_fitter(data, prior, bounds)` takes them (ttv_fit.py:15,
    :199-262): data[i] = {'Tc': ..., 'Tc_err': ...} or {} per body; prior = objects list; bounds[i] =
    {key: [lo, hi]} per body."""

    def __init__(self, data, prior, bounds, device=0):
        self.data, self.prior, self.bounds, self.device = data, copy.deepcopy(prior), bounds, device
        self.freekeys, boundarray = [], []
        for i, planet in enumerate(bounds):                      # ttv_fit.py:26-31
            for bound in planet:
                self.freekeys.append((i, bound))
                boundarray.append(planet[bound])
        self.boundarray = np.array(boundarray, dtype=float)
        tc1 = np.asarray(data[1]['Tc'], dtype=float)
        self.sim_time = tc1.max() - tc1.min() + self.prior[1]['P'] * 5      # :34-36
        self.orbit = np.rint((tc1 - tc1.min()) / self.prior[1]['P']).astype(np.int32)   # :38-39
        Np, n_obs = len(prior) - 1, len(self.orbit)
        self.has = np.zeros(Np, dtype=np.int32)
        self.tc, self.err = np.zeros((Np, n_obs)), np.ones((Np, n_obs))
        for i in range(1, len(prior)):
            if data[i]:
                if len(data[i]['Tc']) != n_obs:
                    raise ValueError("every planet's data must have the length of data[1]['Tc'] (Tc_sim[self.orbit])")
                self.has[i - 1] = 1
                self.tc[i - 1], self.err[i - 1] = data[i]['Tc'], data[i]['Tc_err']

    def prior_transform(self, upars):                            # ttv_fit.py:96-97
        return self.boundarray[:, 0] + np.diff(self.boundarray, 1).reshape(-1) * upars

    def params_for(self, thetas):
        """[B, n_bodies, 7] parameter block for B sampler vectors (the mapping of :49-56)."""
        thetas = np.atleast_2d(np.asarray(thetas, dtype=float))
        base = _abi.objects_to_params(self.prior)
        par = np.repeat(base[None], len(thetas), axis=0)
        cols = {"m": _abi.P_M, "P": _abi.P_P, "e": _abi.P_E, "inc": _abi.P_INC, "Omega": _abi.P_OMEGA_NODE,
                "omega": _abi.P_OMEGA_PERI, "f": _abi.P_F}
        for c, (idx, key) in enumerate(self.freekeys):
            if key == 'tmid':
                continue
            par[:, idx, cols[key]] = thetas[:, c]
        return par

    def loglike_batch(self, thetas, fused=True):
        """chi2 of ttv_fit.py:58-94 for every row of `thetas`.  fused (default): one nbt_run with the likelihood chained
        behind the forward model on the device — only loglike[B] comes back; fused=False keeps the two-call form
        (nbt_run, then nbt_ttvfit_loglike on the returned transit times)."""
        lib = _abi.load()
        par = self.params_for(thetas)
        if fused:
            like = engine.Likelihood("ttvfit", has_data=self.has, tc=self.tc, tc_err=self.err, orbit=self.orbit)
            plan = engine.Plan(par, self.sim_time, int(self.sim_time * 24), flags=0, device=self.device, likelihood=like)   # :59
            buf = engine.Buffers(plan, skip=("tt", "ttv", "n_tt", "period", "t0", "ttv_max"))
            engine.run(plan, buf)
            return buf.loglike.copy()
        plan = engine.Plan(par, self.sim_time, int(self.sim_time * 24), flags=0, device=self.device)   # :59
        buf = engine.run(plan)
        out = np.zeros(plan.B)
        _abi.check(lib.nbt_ttvfit_loglike(_abi.ptr(buf.tt), _abi.ptr(plan.tt_offsets, C.c_int64), _abi.ptr(buf.n_tt, C.c_int32),
                                          plan.B, plan.Np, _abi.ptr(self.has, C.c_int32), _abi.ptr(self.tc), _abi.ptr(self.err),
                                          _abi.ptr(self.orbit, C.c_int32), len(self.orbit), _abi.ptr(out), self.device, 0, None), lib)
        return out

    def loglike(self, pars):
        return float(self.loglike_batch(np.asarray(pars)[None])[0])
