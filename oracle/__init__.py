"""CPU ORACLE — test infrastructure, not product code.

ctypes front-end of oracle/_build/libnbt_oracle.so (built from oracle/nbt_oracle.c by
oracle/Makefile).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this package.  See nbt_oracle.c's header for what is restated
and how parity is pinned.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from nbody_ai_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnbt_oracle.so")
_lib = None
_dp, _ip, _lp = _abi._dp, _abi._ip, _abi._lp


def build(force=False):
    src = os.path.join(_HERE, "nbt_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.nbo_initial_state.argtypes = [_dp, C.c_int, _dp]
        L.nbo_integrate_series.restype = C.c_long
        L.nbo_integrate_series.argtypes = [_dp, C.c_int, C.c_double, C.c_int, _dp, _dp]
        L.nbo_integrate_series_whfast.restype = C.c_long
        L.nbo_integrate_series_whfast.argtypes = [_dp, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int, _dp]
        L.nbo_transit_times.restype = C.c_int
        L.nbo_transit_times.argtypes = [_dp, _dp, _dp, C.c_int, _dp, C.c_int]
        L.nbo_ttv_fit.argtypes = [_dp, C.c_int, _dp, _dp, _dp]
        L.nbo_maxavg.restype = C.c_double
        L.nbo_maxavg.argtypes = [_dp, C.c_int]
        L.nbo_ls_nfreq.restype = C.c_int64
        L.nbo_ls_nfreq.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
        L.nbo_lomb_scargle.argtypes = [_dp, _dp, C.c_int, C.c_double, C.c_double, C.c_int64, _dp]
        L.nbo_loglike.restype = C.c_double
        L.nbo_loglike.argtypes = [_dp, C.c_int, C.c_double, C.c_double, _dp, _dp, _dp, C.c_int]
        L.nbo_loglike_loss.restype = C.c_double
        L.nbo_loglike_loss.argtypes = [_dp, C.c_int, C.c_double, C.c_double, _dp, _dp, _dp, C.c_int, C.c_int]
        L.nbo_refined_transit_times.restype = C.c_int
        L.nbo_refined_transit_times.argtypes = [_dp, C.c_int, C.c_double, C.c_int, C.c_int, _dp, C.c_int]
        L.nbo_run.argtypes = [C.POINTER(_abi.nbt_config), C.POINTER(_abi.nbt_inputs),
                              C.POINTER(_abi.nbt_outputs), C.c_int, _lp]
        L.nbo_max_threads.restype = C.c_int
        _lib = L
    return _lib


SERIES_KEYS = ["x", "y", "z", "P", "a", "e", "inc", "Omega", "omega", "M"]  # simulation.py:13-25


def initial_state(params):
    params = np.ascontiguousarray(params, dtype=np.float64)
    st = np.zeros((params.shape[0], 7))
    lib().nbo_initial_state(_abi.ptr(params), params.shape[0], _abi.ptr(st))
    return st


def integrate_series(params, n_days, n_outputs):
    """Oracle of generate()+integrate(): returns (times, series[n_outputs, 3+10*Np], final_state, n_steps)."""
    params = np.ascontiguousarray(params, dtype=np.float64)
    n = params.shape[0]
    series = np.zeros((n_outputs, 3 + 10 * (n - 1)))
    final = np.zeros((n, 6))
    ns = lib().nbo_integrate_series(_abi.ptr(params), n, float(n_days), int(n_outputs), _abi.ptr(series), _abi.ptr(final))
    times = np.linspace(0.0, n_days, n_outputs)
    return times, series, final, ns


def integrate_series_whfast(params, n_days, n_outputs, dt=0.0, corrector=5):
    """Oracle of generate(..., integrator='whfast') + integrate() (nbody/simulation.py:39-42): WHFast at REBOUND's
    default dt = 1e-3 d (or `dt`), symplectic corrector of order 5 (0 = off).  Returns (times, series, n_steps)."""
    params = np.ascontiguousarray(params, dtype=np.float64)
    n = params.shape[0]
    series = np.zeros((n_outputs, 3 + 10 * (n - 1)))
    ns = lib().nbo_integrate_series_whfast(_abi.ptr(params), n, float(n_days), int(n_outputs), float(dt), int(corrector), _abi.ptr(series))
    return np.linspace(0.0, n_days, n_outputs), series, ns


def transit_times(xp, xs, times):
    xp, xs, times = (np.ascontiguousarray(a, dtype=np.float64) for a in (xp, xs, times))
    tt = np.zeros(len(times))
    n = lib().nbo_transit_times(_abi.ptr(xp), _abi.ptr(xs), _abi.ptr(times), len(times), _abi.ptr(tt), len(tt))
    return tt[:n].copy()


def ttv_fit(tt):
    tt = np.ascontiguousarray(tt, dtype=np.float64)
    ttv = np.zeros(len(tt))
    m, b = C.c_double(), C.c_double()
    lib().nbo_ttv_fit(_abi.ptr(tt), len(tt), _abi.ptr(ttv), C.byref(m), C.byref(b))
    return ttv, m.value, b.value


def maxavg(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    return lib().nbo_maxavg(_abi.ptr(x), len(x))


def lomb_scargle(t, y, minfreq=1.0 / 365, maxfreq=0.5, samples_per_peak=10):
    t = np.ascontiguousarray(t, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    T = t.max() - t.min()
    nf = lib().nbo_ls_nfreq(T, minfreq, maxfreq, samples_per_peak)
    df = 1.0 / T / samples_per_peak
    power = np.zeros(nf)
    lib().nbo_lomb_scargle(_abi.ptr(t), _abi.ptr(y), len(t), minfreq, df, nf, _abi.ptr(power))
    return minfreq + df * np.arange(nf), power


def loglike(ttv, c0, c1, x, y, yerr, loss="linear"):
    ttv, x, y, yerr = (np.ascontiguousarray(a, dtype=np.float64) for a in (ttv, x, y, yerr))
    return lib().nbo_loglike_loss(_abi.ptr(ttv), len(ttv), c0, c1, _abi.ptr(x), _abi.ptr(y), _abi.ptr(yerr), len(x),
                                  _abi.LOSSES[loss])


def refined_transit_times(params, n_days, n_outputs, planet):
    """True roots of x_planet - x_star on the IAS15 trajectory (bracketed on the output grid, Newton-refined)."""
    params = np.ascontiguousarray(params, dtype=np.float64)
    out = np.zeros(int(n_outputs))
    n = lib().nbo_refined_transit_times(_abi.ptr(params), params.shape[0], float(n_days), int(n_outputs), int(planet),
                                        _abi.ptr(out), len(out))
    return out[:n].copy()


def run(cfg, params, buffers, n_threads=0):
    """Batched oracle with the product's buffer layout (nbody_ai_b200.engine.Buffers)."""
    inp, out = buffers.as_structs(params)
    steps = C.c_int64(0)
    rc = lib().nbo_run(C.byref(cfg), C.byref(inp), C.byref(out), int(n_threads), C.byref(steps))
    assert rc == 0
    return steps.value


def max_threads():
    return lib().nbo_max_threads()
