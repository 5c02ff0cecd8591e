"""Drop-in for the training-set generator simulations/generate_simulations.py (:41-110): random
2-planet systems -> planet-1 transit times, with the reference's rejection rule and omega sweep,
written in the reference's pickle format [X, y] so that simulations/fourier_*.py and
encoder_single.py consume GPU-generated data unchanged.  The per-sample Python loop
(randomize -> generate -> analyze, 1 + `omegas` simulations each) becomes batched nbt_run calls."""
import copy
import pickle

import numpy as np

from . import _abi, engine
from .simulation import randomize
from .tools import mearth, msun


def _xn(objects):
    """The 7-vector stored per simulation (generate_simulations.py:68-76)."""
    return [objects[0]['m'], objects[1]['P'], objects[1]['m'] * msun / mearth, objects[2]['P'],
            objects[2]['m'] * msun / mearth, objects[2]['omega'], objects[2]['e']]


def _run(object_lists, periods, device):
    """Planet-1 transit products of each system on the reference's own output grid,
    np.linspace(0, P1*periods, round(P1*periods*24)) (generate_simulations.py:47-50).  The grid
    depends on P1, so systems are grouped by (Ndays, Noutputs) — the omega variants of one draw share
    a grid and go out as one nbt_run — and returned as (tt, maxavg(ttv), n_tt) per system."""
    groups = {}
    for i, ob in enumerate(object_lists):
        P1 = ob[1]['P']
        groups.setdefault((P1 * periods, int(round(P1 * periods * 24))), []).append(i)
    results = [None] * len(object_lists)
    for (ndays, nout), idx in groups.items():
        par = np.stack([_abi.objects_to_params(object_lists[i]) for i in idx])
        plan = engine.Plan(par, ndays, nout, flags=_abi.F_PLANET1_ONLY, device=device)
        buf = engine.run(plan)
        for q, i in enumerate(idx):
            results[i] = (buf.tt_of(q, 0).copy(), float(buf.ttv_max[q * plan.Np]), int(buf.n_tt[q * plan.Np]))
    return results


def generate_training_set(samples, omegas=6, periods=1000, file=None, rng=None, device=0, X=None, y=None,
                          dump_every=100):
    """generate_simulations.py:41-110.  Returns (X, y): X rows are the 7-vectors, y the first
    `periods` planet-1 mid-transit times.  A draw is re-sampled while it has fewer than `periods`
    transits or a TTV amplitude (maxavg) above 180 min (:54-55); every accepted draw is followed by
    `omegas` variants with omega_2 = linspace(-pi, pi, omegas) + N(0, pi/omegas) (:83-90).
    X, y continue an existing set (checkpoint/resume, :27-30); `file` is dumped every `dump_every`
    samples and at the end (:109-110)."""
    R = np.random if rng is None else rng
    X = [] if X is None else X
    y = [] if y is None else y
    start = len(X)
    for ii in range(start, samples):
        while True:
            og = randomize(rng)
            tt, ttv_max, n = _run([og], periods, device)[0]
            if not (n < periods or ttv_max * 24 * 60 > 180):
                break
        X.append(_xn(og))
        y.append(tt[:periods])
        variants = []
        for j in np.linspace(-np.pi, np.pi, omegas) + R.normal(0, np.pi / omegas, omegas):
            ob = copy.deepcopy(og)
            ob[2]['omega'] = j
            variants.append(ob)
        for ob, (tt, _, _) in zip(variants, _run(variants, periods, device)):
            X.append(_xn(ob))
            y.append(tt[:periods])
        if file is not None and ii % dump_every == 0:
            pickle.dump([X, y], open(file, 'wb'))
    if file is not None:
        pickle.dump([X, y], open(file, 'wb'))
    return X, y
