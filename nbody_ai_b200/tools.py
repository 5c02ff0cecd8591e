"""Drop-in for nbody/tools.py of pearsonkyle/Nbody-AI (same names, arguments, return shapes).

Constants and the scalar prior helpers (sa, rv_semi, hill_sphere, m2r_star: nbody/tools.py:11-59)
are host-side input generation, as in the reference.  Everything that is arithmetic on the hot
path — find_zero/transit interpolation, TTV, maxavg, lomb_scargle — is evaluated by libnbt.so
on the GPU through the C ABI (include/nbt.h); there is no CPU fallback, the calls raise when the
CUDA library or a device is missing.
"""
import ctypes as C

import numpy as np

from . import _abi

rsun = 6.955e8      # m            nbody/tools.py:11-18
msun = 1.989e30     # kg
mjup = 1.898e27     # kg
rjup = 7.1492e7     # m
mearth = 5.972e24   # kg
rearth = 6.3781e6   # m
au = 1.496e11       # m
G = 0.00029591220828559104  # day, AU, Msun


def stellar_mass(logg, rs):  # nbody/tools.py:21
    return ((rs * 6.955e10) ** 2 * 10 ** logg / 6.674e-8) / 1000. / msun


def sa(m, P):  # nbody/tools.py:25 — keplerian semi-major axis (au)
    return (G * m * P ** 2 / (4 * np.pi ** 2)) ** (1. / 3)


def rv_semi(M, m, a):  # nbody/tools.py:28 — approximate radial velocity semi-amplitude
    return (G * m * m / (M * a)) ** 0.5


# stellar mass -> radius relation (nbody/tools.py:32-36, scipy interp1d == linear interpolation)
_mms = [40, 18, 6.5, 3.2, 2.1, 1.7, 1.3, 1.1, 1, 0.93, 0.78, 0.69, 0.47, 0.21, 0.1]
_rrs = [18, 7.4, 3.8, 2.5, 1.7, 1.3, 1.2, 1.05, 1, 0.93, 0.85, 0.74, 0.63, 0.32, 0.13]


def m2r_star(m):
    m = np.asarray(m, dtype=float)
    if np.any(m < _mms[-1]) or np.any(m > _mms[0]):
        raise ValueError("A value in x_new is outside the interpolation range.")
    return np.interp(m, _mms[::-1], _rrs[::-1])


def hill_sphere(objs, i=1):
    """Hill-sphere radius between the star and the i-th planet (nbody/tools.py:38-59):
    root of the force balance a_star + a_planet - a_centripetal on [0.5, 1.5] x the
    (m/3M)^(1/3) estimate."""
    from scipy.optimize import brentq
    obj = objs[i]
    M = objs[0]['m']
    a = obj['a'] if 'a' in obj else (G * M * obj['P'] ** 2 / (4 * np.pi ** 2)) ** (1. / 3)
    hill_init = a * (obj['m'] / (3 * M)) ** (1. / 3)

    def hill_radius(r):
        a_cent = (G * M / a) / a
        a_star = G * M / (a + r) ** 2
        a_planet = G * obj['m'] / r ** 2
        return a_star + a_planet - a_cent

    return brentq(hill_radius, hill_init * 0.5, hill_init * 1.5)


# ---------------------------------------------------------------------------------------------
# hot-path functions: thin calls into libnbt.so
# ---------------------------------------------------------------------------------------------
def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def find_zero(t1, dx1, t2, dx2):
    """Zero by linear interpolation between two samples (nbody/tools.py:61-65); scalars or arrays."""
    a = [np.atleast_1d(_f64(v)) for v in (t1, dx1, t2, dx2)]
    a = [np.ascontiguousarray(v) for v in np.broadcast_arrays(*a)]
    lib = _abi.load()
    out = np.zeros(a[0].shape)
    _abi.check(lib.nbt_find_zero(_abi.ptr(a[0]), _abi.ptr(a[1]), _abi.ptr(a[2]), _abi.ptr(a[3]), a[0].size,
                                 _abi.ptr(out), 0, 0, None), lib)
    return float(out[0]) if np.ndim(t1) == 0 and np.ndim(dx1) == 0 else out


def transit_times(xp, xs, times):
    """Mid-transit times of one planet (nbody/simulation.py:163-170): +x -> -x crossings of
    xp - xs, linearly interpolated on the sample grid."""
    xp, xs, times = _f64(xp), _f64(xs), _f64(times)
    n = len(times)
    if not (len(xp) == len(xs) == n):
        raise ValueError("xp, xs, times must have equal length")
    if n < 2:
        return np.zeros(0)
    lib = _abi.load()
    tt = np.zeros(n - 1)
    cnt = np.zeros(1, dtype=np.int32)
    _abi.check(lib.nbt_transit_times(_abi.ptr(xp), _abi.ptr(xs), _abi.ptr(times), 1, n, _abi.ptr(tt),
                                     _abi.ptr(cnt, C.c_int32), n - 1, 1, 0, 0, None), lib)
    return tt[:int(cnt[0])].copy()


def TTV(epochs, tt):
    """Linear ephemeris fit (nbody/tools.py:67-76; shadowed by nbody/simulation.py:172-177):
    returns [ttv, m, b] with tt ~ b + m*epoch."""
    ep, t = _f64(epochs), _f64(tt)
    if len(ep) != len(t):
        raise ValueError("epochs and tt must have equal length")
    if len(t) == 0:
        raise np.linalg.LinAlgError("TTV: empty input")  # np.linalg.lstsq raises on 0 rows
    lib = _abi.load()
    ttv = np.zeros(len(t))
    mb = np.zeros(2)
    _abi.check(lib.nbt_ttv_fit(_abi.ptr(ep), _abi.ptr(t), 1, len(t), _abi.ptr(ttv), _abi.ptr(mb[0:1]), _abi.ptr(mb[1:2]),
                               0, 0, None), lib)
    return [ttv, float(mb[0]), float(mb[1])]


def maxavg(x):
    """(percentile(|x|, 75) + max|x|) / 2 — nbody/tools.py:22."""
    x = _f64(x).ravel()
    if len(x) == 0:
        raise ValueError("maxavg of an empty array")
    lib = _abi.load()
    out = np.zeros(1)
    _abi.check(lib.nbt_maxavg(_abi.ptr(x), 1, len(x), _abi.ptr(out), 0, 0, None), lib)
    return float(out[0])


def make_sin(t, pars, freqs=-1):
    """Sum of sines (nbody/tools.py:78-95): rows of pars are (frequency, amplitude, phase), or (amplitude, phase)
    with the frequencies given in `freqs`.  Host helper of lomb_scargle(npeaks > 0) and of the decoder scripts."""
    fn = 0
    pars = np.asarray(pars, dtype=float)
    if isinstance(freqs, (list, np.ndarray)):
        if pars.ndim == 1:
            pars = pars.reshape(-1, 2)
        for i in range(pars.shape[0]):
            fn += pars[i, 0] * np.sin(t * 2 * np.pi * freqs[i] + pars[i, 1])
    else:
        if pars.ndim == 1:
            pars = pars.reshape(-1, 3)
        for i in range(pars.shape[0]):
            fn += pars[i, 1] * np.sin(t * 2 * np.pi * pars[i, 0] + pars[i, 2])
    return fn


def _weighted_ls_host(t, y, dy, freq):
    """Floating-mean generalised Lomb-Scargle with weights 1/dy^2 (LombScargle(t, y, dy), standard normalisation)
    [3P].  Never on the hot path (no caller of the reference passes dy: nbody/simulation.py:188,213), so it is a
    plain numpy evaluation on the host, kept for signature compatibility."""
    w = 1.0 / dy ** 2
    w = w / w.sum()
    yb = np.sum(w * y)
    yc = y - yb
    YY = np.sum(w * yc * yc)
    power = np.zeros(len(freq))
    for lo in range(0, len(freq), 256):
        f = freq[lo:lo + 256, None]
        arg = 2 * np.pi * f * t[None, :]
        s, c = np.sin(arg), np.cos(arg)
        S, Cc = (w * s).sum(1), (w * c).sum(1)
        S2 = 2 * (w * s * c).sum(1) - 2 * S * Cc
        C2 = (w * (c * c - s * s)).sum(1) - (Cc * Cc - S * S)
        tau = 0.5 * np.arctan2(S2, C2)
        ct, st = np.cos(tau)[:, None], np.sin(tau)[:, None]
        cs, sn = c * ct + s * st, s * ct - c * st
        YC, YS = (w * yc * cs).sum(1), (w * yc * sn).sum(1)
        CC = (w * cs * cs).sum(1) - (w * cs).sum(1) ** 2
        SS = (w * sn * sn).sum(1) - (w * sn).sum(1) ** 2
        with np.errstate(divide="ignore", invalid="ignore"):
            power[lo:lo + 256] = (np.where(CC > 1e-14, YC * YC / CC, 0.0) + np.where(SS > 1e-14, YS * YS / SS, 0.0)) / YY
    return power


def lomb_scargle(t, y, dy=None, minfreq=1. / 365, maxfreq=1 / 2, npeaks=0, peaktol=0.05):
    """Periodogram on astropy's autofrequency grid (nbody/tools.py:97-145): returns (frequency, power), or
    (frequency, power, fdata) when npeaks > 0.

    Floating-mean generalised Lomb-Scargle with standard normalisation (what astropy's exact methods compute), on
    the GPU.  The two off-path branches keep the reference's signature and run on the host: `dy` as an ndarray
    (per-point errors; a non-array dy is ignored like the reference's isinstance test, :99-102) and `npeaks > 0`
    (find_peaks + bounded least squares of a sine series, :112-143; fdata rows = frequency, amplitude, phase)."""
    t, y = _f64(t), _f64(y)
    if len(t) != len(y):
        raise ValueError("t and y must have equal length")
    lib = _abi.load()
    baseline = float(t.max() - t.min())
    nf = int(lib.nbt_ls_nfreq(baseline, float(minfreq), float(maxfreq), 10))
    if nf <= 0:
        raise ValueError("lomb_scargle: zero baseline")
    df = 1.0 / baseline / 10
    frequency = minfreq + df * np.arange(nf)
    if isinstance(dy, np.ndarray):
        power = _weighted_ls_host(t, y, _f64(dy), frequency)
    else:
        power = np.zeros(nf)
        _abi.check(lib.nbt_lomb_scargle(_abi.ptr(t), _abi.ptr(y), 1, len(t), float(minfreq), df, nf, _abi.ptr(power),
                                        0, 0, None), lib)
    if npeaks > 0:
        from scipy.optimize import least_squares
        from scipy.signal import find_peaks
        peaks, amps = find_peaks(power, height=peaktol)
        Nterms = npeaks
        fdata = np.zeros((Nterms, 3))                                     # frequency, amplitude, shift
        peaks = peaks[np.argsort(amps['peak_heights'])[::-1]]             # sort high to low
        for i in range(min(npeaks, len(peaks))):
            fdata[i, 0] = frequency[int(peaks[i])]
            fdata[i, 1] = np.sort(amps['peak_heights'])[::-1][i] * maxavg(y)
            fdata[i, 2] = 0
        priors = fdata.flatten()
        bounds = np.array([[0, 1], [0, max(y) * 1.5], [0, 2 * np.pi]] * Nterms).T

        def fit_wave(pars):
            return (y - make_sin(t, pars)) / y

        res = least_squares(fit_wave, x0=priors, bounds=bounds)
        return frequency, power, res.x.reshape(Nterms, -1)
    return frequency, power
