"""Drop-in for the compute part of simulations/simulation_grid.py (generate_sim :20-69 and the grid
loop :136-153): TTV amplitude / periodicity / harmonic-model grids over perturber mass x period.
The reference evaluates one cell at a time (generate -> transit_times -> TTV -> LombScargle ->
find_peaks -> lstsq); here the whole grid is one nbt_run + one nbt_grid_post on the GPU.
write_fits() lays the grids out like the reference's FITS file (:219-295; written by hand — astropy is not a
dependency); the PNG plots (:155-217) are out of scope."""
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
    per_epoch_grid [n_order, N, N], amplitude_grid [2 n_order, N, N], error_grid, ttv_grid, rv_grid.
    objects[2] needs no 'P' (the reference's list has none until the loop sets it)."""
    mass_grid, period_grid = np.asarray(mass_grid, dtype=float), np.asarray(period_grid, dtype=float)
    shape = mass_grid.shape
    objs = [dict(o) for o in objects]
    objs[2].setdefault('P', float(period_grid.flat[0]))
    base = _abi.objects_to_params(objs)
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


# ---------------------------------------------------------------------------------------------
# FITS layout of simulation_grid.py:219-295: primary HDU = TTV amplitude [min], image extensions = first and second
# O-C periodicity [epochs] and RV amplitude [m/s] (both masked to NaN where the amplitude is <= 1 min, :174-179),
# linear axes in CRVAL/CDELT, the objects in the header, EXTNAMEs.  Plain FITS 4.0: 2880-byte blocks of 80-character
# cards, big-endian IEEE doubles.
# ---------------------------------------------------------------------------------------------
def _card(key, value=None, comment=None):
    if key == "COMMENT":
        return ("COMMENT " + str(value))[:80].ljust(80)
    if isinstance(value, bool):
        v = "T" if value else "F"
        body = "%-8s= %20s" % (key, v)
    elif isinstance(value, (int, np.integer)):
        body = "%-8s= %20d" % (key, value)
    elif isinstance(value, (float, np.floating)):
        txt = repr(float(value)).upper()
        if "." not in txt and "E" not in txt and "N" not in txt:
            txt += ".0"
        body = "%-8s= %20s" % (key, txt)
    else:
        body = "%-8s= %-20s" % (key, "'%-8s'" % str(value).replace("'", "''"))
    if comment:
        body += " / " + comment
    return body[:80].ljust(80)


def _hdu(data, cards, primary):
    data = np.ascontiguousarray(data, dtype=">f8")
    head = [_card("SIMPLE", True, "conforms to FITS standard")] if primary else [_card("XTENSION", "IMAGE", "Image extension")]
    head += [_card("BITPIX", -64, "array data type"), _card("NAXIS", data.ndim, "number of array dimensions")]
    for i, n in enumerate(data.shape[::-1]):
        head.append(_card("NAXIS%d" % (i + 1), int(n)))
    head += [_card("EXTEND", True)] if primary else [_card("PCOUNT", 0, "number of parameters"), _card("GCOUNT", 1, "number of groups")]
    head += [_card(*c) for c in cards] + ["END".ljust(80)]
    raw = "".join(head).encode("ascii")
    raw += b" " * (-len(raw) % 2880)
    body = data.tobytes()
    return raw + body + b"\0" * (-len(body) % 2880)


def write_fits(path, grids, mass_grid, period_grid, objects, msun_over_mearth):
    """ttv_grid.fits of simulation_grid.py:219-295 from the dict simulation_grid() returns.  objects: the list with
    planet 2 as the loop left it (last cell); msun_over_mearth = msun / mearth of the caller's constants (:83-84)."""
    npts = mass_grid.shape[0]
    ttv_min = grids["ttv_grid"] * 24 * 60
    mask = ttv_min > 1                                                         # :174-179
    per1 = np.where(mask, grids["per_epoch_grid"][0], np.nan)
    per2 = np.where(mask, grids["per_epoch_grid"][1], np.nan) if len(grids["per_epoch_grid"]) > 1 else np.full_like(per1, np.nan)
    rv = np.where(mask, grids["rv_grid"], np.nan)
    m_lo, m_hi = mass_grid.min() * msun_over_mearth, mass_grid.max() * msun_over_mearth
    p_lo, p_hi = period_grid.min() / objects[1]['P'], period_grid.max() / objects[1]['P']

    def axes(crpix_first):
        x = [("CRVAL1", float(m_lo)), ("CRPIX1", 1), ("CDELT1", float((m_hi - m_lo) / npts))]
        yy = [("CRVAL2", float(p_lo)), ("CRPIX2", 1), ("CDELT2", float((p_hi - p_lo) / npts))]
        if crpix_first:                                                       # the primary HDU sets CRPIX before CRVAL (:223-230)
            x, yy = [x[1], x[0], x[2]], [yy[1], yy[0], yy[2]]
        return x + yy

    objs = [("STARMASS", float(objects[0]['m']))]
    for i, tag in ((1, "P1"), (2, "P2")):
        objs += [(tag + "_MASS", float(objects[i]['m'])), (tag + "_PER", float(objects[i]['P'])), (tag + "_INC", float(objects[i]['inc'])),
                 (tag + "_ECC", float(objects[i]['e'])), (tag + "_OMEGA", float(objects[i]['omega']))]
    hdus = [
        (ttv_min, [("BUNIT", "minutes"), ("COMMENT", "TTV amplitude in minutes")] + axes(True) + objs + [("EXTNAME", "TTV_AMPLITUDE")], True),
        (per1, [("BUNIT", "epochs"), ("COMMENT", "Periodicity of O-C in epochs")] + axes(False) + objs + [("EXTNAME", "TTV_PERIODICITY1")], False),
        (per2, [("BUNIT", "epochs"), ("COMMENT", "Periodicity of O-C in epochs")] + axes(False) + objs + [("EXTNAME", "TTV_PERIODICITY2")], False),
        (rv, [("BUNIT", "m/s"), ("COMMENT", "Radial velocity amplitude in m/s")] + axes(False) + objs + [("EXTNAME", "RV_AMPLITUDE")], False),
    ]
    with open(path, "wb") as fh:
        for data, cards, primary in hdus:
            fh.write(_hdu(data, cards, primary))


def read_fits(path):
    """Minimal reader of the files write_fits produces (tests / quick looks): list of (header dict, array)."""
    raw = open(path, "rb").read()
    out, pos = [], 0
    while pos < len(raw):
        hdr, end = {}, False
        while not end:
            block = raw[pos:pos + 2880].decode("ascii")
            pos += 2880
            for i in range(0, 2880, 80):
                card = block[i:i + 80]
                key = card[:8].strip()
                if key == "END":
                    end = True
                    break
                if key == "COMMENT":
                    hdr.setdefault("COMMENT", []).append(card[8:].strip())
                elif card[8:10] == "= ":
                    val = card[10:].split(" / ")[0].strip()
                    if val.startswith("'"):
                        hdr[key] = val.strip("'").rstrip()
                    elif val in ("T", "F"):
                        hdr[key] = val == "T"
                    else:
                        hdr[key] = float(val) if any(c in val for c in ".EN") else int(val)
        shape = tuple(hdr["NAXIS%d" % (i + 1)] for i in range(hdr["NAXIS"]))[::-1]
        n = int(np.prod(shape)) * 8
        out.append((hdr, np.frombuffer(raw[pos:pos + n], dtype=">f8").reshape(shape).astype(float)))
        pos += n + (-n % 2880)
    return out
