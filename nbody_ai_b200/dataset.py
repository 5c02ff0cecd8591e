"""Drop-in for the training-set generator simulations/generate_simulations.py (:41-110): random
2-planet systems -> planet-1 transit times, with the reference's rejection rule and omega sweep,
written in the reference's pickle format [X, y] so that simulations/fourier_*.py and
encoder_single.py consume GPU-generated data unchanged.

The reference runs one simulation at a time: randomize() -> generate(Ndays = P1 * periods, Noutputs =
round(P1 * periods * 24)) -> analyze(), redrawn while the draw is bad, then `omegas` more simulations of the
accepted draw.  Every draw has its own output grid, which is why this used to degenerate into one tiny launch
per draw.  Here a whole block of candidate draws AND all their omega variants go out as ONE nbt_run with
per-system grids (nbt_inputs.n_days_sys / n_outputs_sys); the rejection rule (:54-55) is applied to the
results on the host and the variants of rejected candidates are simply dropped (speculative work: the
acceptance rate of the priors is ~85 %)."""
import pickle

import numpy as np

from . import _abi, engine
from .simulation import randomize
from .tools import mearth, msun


def _xn(row):
    """The 7-vector stored per simulation (generate_simulations.py:68-76) from a parameter block [3][7]."""
    return [float(row[0, _abi.P_M]), float(row[1, _abi.P_P]), float(row[1, _abi.P_M] * msun / mearth),
            float(row[2, _abi.P_P]), float(row[2, _abi.P_M] * msun / mearth), float(row[2, _abi.P_OMEGA_PERI]),
            float(row[2, _abi.P_E])]


def grids(params, periods):
    """The reference's output grid of every system: Ndays = P1 * periods, Noutputs = round(P1 * periods * 24)
    (generate_simulations.py:45-50; Python's round = round-half-even = np.rint)."""
    P1 = params[:, 1, _abi.P_P]
    n_days = P1 * periods
    return n_days, np.rint(P1 * periods * 24).astype(np.int32)


def draw_block(n, omegas, rng=None):
    """n candidate draws of randomize() (the reference's RNG consumption, draw by draw) and, for each, the omega sweep
    omega_2 = linspace(-pi, pi, omegas) + N(0, pi / omegas) (:83-90).  Returns params [n, 1 + omegas, 3, 7]: slot 0 is
    the draw itself.  The reference draws the sweep noise only for accepted draws; here every candidate gets one
    (candidates are speculative), so a seeded run follows the reference's priors, not its exact random stream."""
    R = np.random if rng is None else rng
    out = np.zeros((n, 1 + omegas, 3, _abi.NBT_NPAR))
    for i in range(n):
        base = _abi.objects_to_params(randomize(rng))
        out[i] = base
        if omegas:
            out[i, 1:, 2, _abi.P_OMEGA_PERI] = np.linspace(-np.pi, np.pi, omegas) + R.normal(0, np.pi / omegas, omegas)
    return out


def run_block(block, periods, device=0, verified=False):
    """Planet-1 transit times of every system of a block [n, 1 + omegas, 3, 7] in one launch.  Returns the engine
    Buffers (systems in row-major order of the block)."""
    n, v = block.shape[:2]
    par = block.reshape(n * v, 3, _abi.NBT_NPAR)
    n_days, n_out = grids(par, periods)
    plan = engine.Plan(par, n_days, n_out, flags=_abi.F_PLANET1_ONLY, device=device)
    return engine.run_verified(plan) if verified else engine.run(plan)


def accepted(buf, n, v, periods):
    """The reference's rejection rule on the draws (slot 0 of each candidate): kept unless
    len(ttv) < periods or maxavg(ttv) > 180 min (generate_simulations.py:54-55)."""
    Np = buf.plan.Np
    rows = np.arange(n) * v
    n_tt = np.minimum(buf.n_tt[rows * Np], np.diff(buf.plan.tt_offsets)[rows * Np])
    n_ttv = np.where(n_tt >= 3, n_tt, 1)                      # analyze(): < 3 transits -> ttv = [0]
    return ~((n_ttv < periods) | (buf.ttv_max[rows * Np] * 24 * 60 > 180))


def generate_training_set(samples, omegas=6, periods=1000, file=None, rng=None, device=0, X=None, y=None,
                          dump_every=100, block=None, verified=False, stats=None):
    """generate_simulations.py:41-110.  Returns (X, y): X rows are the 7-vectors, y the first
    `periods` planet-1 mid-transit times.  A draw is re-sampled while it has fewer than `periods`
    transits or a TTV amplitude (maxavg) above 180 min (:54-55); every accepted draw is followed by
    `omegas` variants with omega_2 = linspace(-pi, pi, omegas) + N(0, pi/omegas) (:83-90).
    X, y continue an existing set (checkpoint/resume, :27-30); `file` is dumped every `dump_every`
    draws and at the end (:109-110).  block: candidate draws per launch (default: what is still missing, + 25 %)."""
    X = [] if X is None else X
    y = [] if y is None else y
    per_draw = 1 + omegas
    done = len(X) // per_draw
    since_dump = 0
    st = dict(launches=0, candidates=0, accepted=0, simulations=0)
    while done < samples:
        need = samples - done
        n = int(block) if block else max(8, int(np.ceil(need * 1.25)))
        blk = draw_block(n, omegas, rng)
        buf = run_block(blk, periods, device=device, verified=verified)
        keep = accepted(buf, n, per_draw, periods)
        st["launches"] += 1; st["candidates"] += n; st["accepted"] += int(keep.sum()); st["simulations"] += n * per_draw
        for i in np.flatnonzero(keep):
            if done >= samples:
                break
            for q in range(per_draw):
                b = i * per_draw + q
                X.append(_xn(blk[i, q]))
                y.append(buf.tt_of(b, 0)[:periods].copy())
            done += 1
            since_dump += 1
            if file is not None and since_dump >= dump_every:
                pickle.dump([X, y], open(file, 'wb'))
                since_dump = 0
    if file is not None:
        pickle.dump([X, y], open(file, 'wb'))
    if stats is not None:
        stats.update(st)
    return X, y
