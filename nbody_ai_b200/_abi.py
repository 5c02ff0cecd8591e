"""ctypes mirror of include/nbt.h (the C ABI of libnbt.so).

Nothing here computes anything: it declares the structs/enums of the boundary and loads the
CUDA library.  There is deliberately no CPU fallback — if libnbt.so is missing or no CUDA
device is present, every compute call raises.
"""
import ctypes as C
import os

import numpy as np

NBT_ABI_VERSION = 2
NBT_G = 0.00029591220828559104  # nbody/tools.py:18
NBT_MAX_BODIES = 9
NBT_NPAR = 7
P_M, P_P, P_E, P_INC, P_OMEGA_NODE, P_OMEGA_PERI, P_F = range(7)

SCHEME_WH2, SCHEME_Y4, SCHEME_SABA4, SCHEME_M64, SCHEME_C4, SCHEME_SBABC3, SCHEME_AUTO, SCHEME_IAS15, SCHEME_WHFAST = range(9)
SCHEMES = {"wh2": SCHEME_WH2, "y4": SCHEME_Y4, "saba4": SCHEME_SABA4, "m64": SCHEME_M64,
           "c4": SCHEME_C4, "sbabc3": SCHEME_SBABC3, "auto": SCHEME_AUTO, "ias15": SCHEME_IAS15, "whfast": SCHEME_WHFAST}
TT_GRID_LINEAR, TT_REFINED = 0, 1

F_MEANS = 1 << 0
F_MEAN_OMEGA = 1 << 1
F_RV = 1 << 2
F_ORBIT = 1 << 3
F_SERIES = 1 << 4
F_LS_OC = 1 << 5
F_LS_RV = 1 << 6
F_PLANET1_ONLY = 1 << 7
F_DEVICE_PTRS = 1 << 8
F_TIMING = 1 << 9
F_LOGLIKE = 1 << 10
F_LS_PEAKS = 1 << 11
LS_NPEAKS = 3
LIKE_NESTED, LIKE_TTVFIT = 0, 1
LOSS_LINEAR, LOSS_SOFT_L1 = 0, 1
LOSSES = {"linear": LOSS_LINEAR, "soft_l1": LOSS_SOFT_L1}

S_NONFINITE = 1 << 0
S_UNBOUND = 1 << 1
S_KEPLER = 1 << 2
S_TT_OVERFLOW = 1 << 3
S_FEW_TRANSITS = 1 << 4
S_UNCALIBRATED = 1 << 5
S_UNRESOLVED = 1 << 6

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


class nbt_config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("n_systems", C.c_int32),
        ("n_bodies", C.c_int32),
        ("n_outputs", C.c_int32),
        ("n_days", C.c_double),
        ("substeps", C.c_int32),
        ("scheme", C.c_int32),
        ("tt_mode", C.c_int32),
        ("flags", C.c_int32),
        ("device", C.c_int32),
        ("orbit_stride", C.c_int32),
        ("ls_oc_minfreq", C.c_double),
        ("ls_oc_maxfreq", C.c_double),
        ("ls_rv_minfreq", C.c_double),
        ("ls_rv_maxfreq", C.c_double),
        ("ls_samples_per_peak", C.c_int32),
        ("tt_cap_max", C.c_int32),
        ("lanes_systems", C.c_int32),
        ("alt_systems", C.c_int32),
    ]


class nbt_like(C.Structure):
    _fields_ = [("kind", C.c_int32), ("loss", C.c_int32), ("n_data", C.c_int32), ("reserved", C.c_int32),
                ("c0", _dp), ("c1", _dp), ("x", _dp), ("y", _dp), ("yerr", _dp),
                ("has_data", _ip), ("tc", _dp), ("tc_err", _dp), ("orbit", _ip)]


class nbt_inputs(C.Structure):
    _fields_ = [("params", _dp), ("substeps", _ip), ("order", _ip), ("tt_offsets", _lp), ("ls_offsets", _lp),
                ("n_days_sys", _dp), ("n_outputs_sys", _ip), ("like", C.POINTER(nbt_like))]


class nbt_outputs(C.Structure):
    _fields_ = [
        ("tt", _dp), ("ttv", _dp), ("n_tt", _ip), ("period", _dp), ("t0", _dp), ("ttv_max", _dp),
        ("mean_a", _dp), ("mean_e", _dp), ("mean_inc", _dp), ("mean_omega", _dp), ("status", _ip),
        ("rv", _dp), ("rv_max", _dp), ("orbit_xz", _dp), ("series", _dp),
        ("ls_oc_power", _dp), ("ls_oc_nfreq", _ip), ("ls_rv_power", _dp), ("ls_oc_peaks", _dp), ("loglike", _dp),
    ]


def make_config(n_systems, n_bodies, n_days, n_outputs, substeps=0, scheme=SCHEME_AUTO,
                tt_mode=TT_GRID_LINEAR, flags=F_MEANS, device=0):
    cfg = nbt_config()
    cfg.struct_size = C.sizeof(nbt_config)
    cfg.n_systems, cfg.n_bodies, cfg.n_outputs = int(n_systems), int(n_bodies), int(n_outputs)
    cfg.n_days = float(n_days)
    cfg.substeps, cfg.scheme, cfg.tt_mode, cfg.flags, cfg.device = (
        int(substeps), int(scheme), int(tt_mode), int(flags), int(device))
    cfg.orbit_stride = 4  # simulation.py:227-228
    cfg.ls_oc_minfreq, cfg.ls_oc_maxfreq = 1.0 / 180, 1.0 / 2  # simulation.py:213
    cfg.ls_rv_minfreq, cfg.ls_rv_maxfreq = 1.0 / 365, 1.0 / 2  # tools.py:97
    cfg.ls_samples_per_peak = 10  # tools.py:104
    return cfg


def ptr(a, ctype=C.c_double):
    """Pointer to a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return C.POINTER(ctype)()
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(ctype))


_LIB = None
# NBT_LIB_PATH: developer override used by tools/gpu_probe.py to compare kernel build variants
_LIB_PATH = os.environ.get("NBT_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnbt.so")


def lib_path():
    return _LIB_PATH


def load():
    """Load libnbt.so (built in-tree by __graft_entry__.build()).  Raises if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            "nbody_ai_b200: %s not found — build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % _LIB_PATH)
    lib = C.CDLL(_LIB_PATH)
    lib.nbt_abi_version.restype = C.c_int
    lib.nbt_device_count.restype = C.c_int
    lib.nbt_last_error.restype = C.c_char_p
    lib.nbt_ls_nfreq.restype = C.c_int64
    lib.nbt_ls_nfreq.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
    lib.nbt_plan_tt.argtypes = [C.POINTER(nbt_config), C.POINTER(nbt_inputs), _lp]
    lib.nbt_plan_order.argtypes = [C.POINTER(nbt_config), C.POINTER(nbt_inputs), _ip]
    lib.nbt_plan_ls.argtypes = [C.POINTER(nbt_config), _lp, _lp]
    lib.nbt_plan_lanes.argtypes = [C.POINTER(nbt_config), C.POINTER(nbt_inputs)]
    lib.nbt_plan_substeps.argtypes = [C.POINTER(nbt_config), C.POINTER(nbt_inputs), _ip]
    lib.nbt_count_steps.restype = C.c_int64
    lib.nbt_count_steps.argtypes = [C.POINTER(nbt_config), C.POINTER(nbt_inputs)]
    lib.nbt_uncalibrated.argtypes = [_dp, C.c_int]
    lib.nbt_plan.argtypes = [C.POINTER(nbt_config), C.POINTER(nbt_inputs), C.c_int32, _ip, _ip, _lp, _lp, _lp, C.c_int32]
    lib.nbt_run.argtypes = [C.POINTER(nbt_config), C.POINTER(nbt_inputs), C.POINTER(nbt_outputs), C.c_void_p]
    lib.nbt_lomb_scargle.argtypes = [_dp, _dp, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_int64,
                                     _dp, C.c_int, C.c_int, C.c_void_p]
    lib.nbt_chi2.argtypes = [_dp, _lp, _ip, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp, _dp, C.c_int32, C.c_int32, _dp,
                             C.c_int, C.c_int, C.c_void_p]
    lib.nbt_analyze.argtypes = [C.POINTER(nbt_config), _dp, _dp, C.c_double, C.POINTER(nbt_inputs),
                                C.POINTER(nbt_outputs), C.c_void_p]
    lib.nbt_transit_times.argtypes = [_dp, _dp, _dp, C.c_int64, C.c_int64, _dp, _ip, C.c_int64, C.c_int, C.c_int,
                                      C.c_int, C.c_void_p]
    lib.nbt_find_zero.argtypes = [_dp, _dp, _dp, _dp, C.c_int64, _dp, C.c_int, C.c_int, C.c_void_p]
    lib.nbt_ttv_fit.argtypes = [_dp, _dp, C.c_int64, C.c_int64, _dp, _dp, _dp, C.c_int, C.c_int, C.c_void_p]
    lib.nbt_maxavg.argtypes = [_dp, C.c_int64, C.c_int64, _dp, C.c_int, C.c_int, C.c_void_p]
    lib.nbt_fp64_peak.argtypes = [C.c_int, C.c_int, _dp, _dp]
    lib.nbt_timing.argtypes = [_dp]
    lib.nbt_ttvfit_loglike.argtypes = [_dp, _lp, _ip, C.c_int32, C.c_int32, _ip, _dp, _dp, _ip, C.c_int32, _dp,
                                       C.c_int, C.c_int, C.c_void_p]
    lib.nbt_grid_post.argtypes = [_dp, _dp, _lp, _ip, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32,
                                  C.c_double, C.c_int32, _dp, _dp, _ip, _dp, _dp, C.c_int, C.c_int, C.c_void_p]
    lib.nbt_fp64_latency.argtypes = [C.c_int, _dp, _dp]
    if lib.nbt_abi_version() != NBT_ABI_VERSION:
        raise RuntimeError("libnbt.so ABI version mismatch")
    _LIB = lib
    return lib


def check(rc, lib=None):
    if rc != 0:
        lib = lib or load()
        msg = lib.nbt_last_error()
        raise RuntimeError("libnbt: rc=%d %s" % (rc, msg.decode() if msg else ""))


def objects_to_params(objects):
    """objects list-of-dicts (README.md:23-28) -> [n_bodies, NBT_NPAR] row block.

    Keys are the ones the reference forwards to rebound.Simulation.add(**obj): m, P (or a), e, inc,
    Omega, omega, f; absent keys default to 0 like REBOUND's.  A semi-major axis `a` is converted to
    the period with the Jacobi mu = G (M_interior + m) that rebound.add uses."""
    out = np.zeros((len(objects), NBT_NPAR))
    keys = {"m": P_M, "P": P_P, "e": P_E, "inc": P_INC, "Omega": P_OMEGA_NODE, "omega": P_OMEGA_PERI, "f": P_F}
    for i, ob in enumerate(objects):
        for k, v in ob.items():
            if k == "a":
                continue
            if k not in keys:
                raise ValueError("unsupported orbital key %r (supported: %s)" % (k, sorted(list(keys) + ["a"])))
            out[i, keys[k]] = float(v)
        if i > 0 and "P" not in ob:
            if "a" not in ob:
                raise ValueError("planet %d needs a period 'P' (or a semi-major axis 'a')" % i)
            mu = NBT_G * out[:i + 1, P_M].sum()
            out[i, P_P] = 2 * np.pi * np.sqrt(float(ob["a"]) ** 3 / mu)
        elif i > 0 and "a" in ob:
            raise ValueError("planet %d: give either 'P' or 'a', not both (rebound.add raises too)" % i)
    return out
