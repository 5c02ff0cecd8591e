"""Drop-in for the compute part of simulations/simulation_grid.py (generate_sim :20-69 and the grid
loop :136-153): TTV amplitude / periodicity / harmonic-model grids over perturber mass x period.
The reference evaluates one cell at a time (generate -> transit_times -> TTV -> LombScargle ->
find_peaks -> lstsq); here the whole grid is one nbt_run + one nbt_grid_post on the GPU.
Plotting and the FITS writer (:155-296) are out of scope."""
import ctypes as C

import numpy as np

from . import _abi, engine
from .simulation import systems_to_params
from .tools import au, rv_semi, sa


def softmax(x):
    """simulation_grid.py:17 — half the 95th percentile of |x| (host helper for callers' own arrays)."""
    return np.percentile(np.abs(x), 95) * 0.5


def generate_sim_batch(systems, epochs=100, n_order=4, peak_height=0.05, device=0, **plan_kw):
    """generate_sim for B systems that share the inner period P1 (a grid does): returns a dict of
    arrays — peak_periods [B, n_order] (0-padded like per_epoch_grid), n_peaks [B],
    coeffs [B, 2 + 2 n_order] (0-padded like amplitude_grid), ttv_max = softmax(ttv) [B], avg_err [B],
    plus the engine buffers under 'buffers'."""
    lib = _abi.load()
    params = systems_to_params(systems)
    P1 = params[:, 1, _abi.P_P]
    if not np.all(P1 == P1[0]):
        raise ValueError("generate_sim_batch: all systems must share P1 (Ndays = P1 * epochs); split the batch")
    ndays = float(P1[0]) * epochs
    nout = int(round(float(P1[0]) * epochs * 48))          # 30-minute cadence (simulation_grid.py:23-25)
    plan = engine.Plan(params, ndays, nout, flags=_abi.F_PLANET1_ONLY, device=device, **plan_kw)
    buf = engine.run(plan)
    B = plan.B
    m = 2 + 2 * n_order
    out = {"peak_periods": np.zeros((B, n_order)), "coeffs": np.zeros((B, m)), "n_peaks": np.zeros(B, dtype=np.int32),
           "ttv_max": np.zeros(B), "avg_err": np.zeros(B), "buffers": buf}
    _abi.check(lib.nbt_grid_post(_abi.ptr(buf.tt), _abi.ptr(buf.ttv), _abi.ptr(plan.tt_offsets, C.c_int64),
                                 _abi.ptr(buf.n_tt, C.c_int32), B, plan.Np, plan.cfg.tt_cap_max,
                                 1. / (1.5 * epochs), 1. / 2.1, 20, float(peak_height), int(n_order),
                                 _abi.ptr(out["peak_periods"]), _abi.ptr(out["coeffs"]), _abi.ptr(out["n_peaks"], C.c_int32),
                                 _abi.ptr(out["ttv_max"]), _abi.ptr(out["avg_err"]), device, 0, None), lib)
    return out


def generate_sim(objects, epochs=100, n_order=4):
    """simulation_grid.py:20-69 for one system: (peak_periods, coeffs, softmax(ttv), avg_err) with the
    reference's shapes (only the peaks that were found; 2 + 2 * len(peaks) coefficients)."""
    r = generate_sim_batch([objects], epochs=epochs, n_order=n_order)
    k = int(r["n_peaks"][0])
    return r["peak_periods"][0, :k].copy(), r["coeffs"][0, :2 + 2 * k].copy(), float(r["ttv_max"][0]), float(r["avg_err"][0])


def simulation_grid(objects, mass_grid, period_grid, epochs=100, n_order=4, device=0):
    """The double loop of simulation_grid.py:136-153 as one batched call.  mass_grid, period_grid:
    [Npts, Npts] arrays of planet-2 mass [Msun] and period [d].  Returns the reference's grids:
    per_epoch_grid [n_order, N, N], amplitude_grid [2 n_order, N, N], error_grid, ttv_grid, rv_grid."""
    mass_grid, period_grid = np.asarray(mass_grid, dtype=float), np.asarray(period_grid, dtype=float)
    shape = mass_grid.shape
    base = _abi.objects_to_params(objects)
    par = np.repeat(base[None], mass_grid.size, axis=0)
    par[:, 2, _abi.P_M] = mass_grid.ravel()
    par[:, 2, _abi.P_P] = period_grid.ravel()
    r = generate_sim_batch(par, epochs=epochs, n_order=n_order, device=device)
    rv = rv_semi(objects[0]['m'], mass_grid, sa(objects[0]['m'], period_grid)) * au / (24 * 60 * 60)
    return {
        "per_epoch_grid": np.moveaxis(r["peak_periods"].reshape(shape + (n_order,)), -1, 0),
        "amplitude_grid": np.moveaxis(r["coeffs"][:, 2:].reshape(shape + (2 * n_order,)), -1, 0),
        "error_grid": r["avg_err"].reshape(shape),
        "ttv_grid": r["ttv_max"].reshape(shape),
        "rv_grid": rv,
    }
