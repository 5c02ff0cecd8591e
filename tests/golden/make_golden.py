"""Generate the committed golden fixtures under tests/golden/.  Run HERE (the container that has
/root/reference); the GPU box only reads the fixtures.

  ref_functions.npz  outputs of the REFERENCE's own pure-numpy functions, imported from
                     /root/reference/nbody with rebound / astropy / matplotlib stubbed out (they are
                     not installable here and are only needed by generate()/lomb_scargle()):
                     transit_times, TTV, find_zero, maxavg, sa, hill_sphere, randomize — fed with
                     series produced by the oracle.  These pin the oracle's restatement of those
                     functions to the reference's actual code.
  sim_data_g2.json   the 30 noisy planet-1 transit times of /root/reference/sim_data.txt (G2) + header
  readme_g1.npz      oracle outputs for the README example (G1), the known-answer vector the CUDA
                     path is compared with on the GPU box
  c2_small.npz       oracle outputs for 48 seeded C2 systems (60 d) — parity fixture
"""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"

import oracle  # noqa: E402
from nbody_ai_b200 import _abi, engine, workloads  # noqa: E402


def import_reference():
    """Import /root/reference/nbody with the missing third-party modules stubbed."""
    for name in ("rebound", "astropy", "astropy.stats", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["astropy.stats"].LombScargle = object
    sys.modules["astropy"].stats = sys.modules["astropy.stats"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    import nbody.simulation as rs
    import nbody.tools as rt
    return rs, rt


def main():
    os.makedirs(OUT, exist_ok=True)
    rs, rt = import_reference()

    # ---- README example through the oracle ------------------------------------------------
    P = workloads.c1_params(1)[0]
    times, series, final, ns = oracle.integrate_series(P, 365, 8760)
    xs = series[:, 0]
    g1 = {"params": P, "initial_state": oracle.initial_state(P), "final_state": final}
    ref = {"times": times, "xs": xs}
    for j in range(3):
        col = series[:, 3 + 10 * j: 13 + 10 * j]
        xp = col[:, 0]
        # the REFERENCE's functions on the oracle's series
        tt_ref = rs.transit_times(xp, xs, times)
        ttv_ref, m_ref, b_ref = rs.TTV(np.arange(len(tt_ref)), tt_ref)
        ref["xp%d" % j] = xp
        ref["tt%d" % j] = tt_ref
        ref["ttv%d" % j] = ttv_ref
        ref["mb%d" % j] = np.array([m_ref, b_ref])
        ref["maxavg%d" % j] = np.array(rt.maxavg(ttv_ref))
        # oracle products
        tt = oracle.transit_times(xp, xs, times)
        ttv, m, b = oracle.ttv_fit(tt)
        f, pw = oracle.lomb_scargle(np.arange(len(ttv)), ttv, 1. / 180, 0.5)
        g1["tt%d" % j], g1["ttv%d" % j], g1["mb%d" % j] = tt, ttv, np.array([m, b])
        g1["means%d" % j] = np.array([col[:, 4].mean(), col[:, 5].mean(), col[:, 6].mean(), col[:, 8].mean()])
        g1["maxavg%d" % j] = np.array(oracle.maxavg(ttv))
        g1["ls_freq%d" % j], g1["ls_power%d" % j] = f, pw
    rv = np.diff(xs) * 1.496e11 * (8760 / (365 * 24 * 60 * 60))
    g1["rv"] = rv
    g1["rv_max"] = np.array(oracle.maxavg(rv))
    f, pw = oracle.lomb_scargle(times[1:], rv)
    g1["rv_ls_freq"], g1["rv_ls_power"] = f, pw
    ref["rv_maxavg"] = np.array(rt.maxavg(rv))
    # find_zero / sa / hill_sphere / randomize from the reference
    ref["find_zero"] = np.array([rt.find_zero(1.0, 0.3, 1.5, -0.2), rt.find_zero(0.0, 1e-3, 1 / 24, -2e-3)])
    ref["sa"] = np.array([rt.sa(1.12, 3.2888), rt.sa(0.8, 20.0)])
    objs = workloads.c1_objects()
    ref["hill"] = np.array([rt.hill_sphere(objs, i=1), rt.hill_sphere(objs, i=2), rt.hill_sphere(objs, i=3)])
    np.random.seed(1234)
    draws = []
    for _ in range(16):
        o = rs.randomize()
        draws.append([o[0]['m'], o[1]['m'], o[1]['P'], o[2]['m'], o[2]['P'], o[2]['e'], o[2]['omega'], o[1]['inc'], o[2]['inc']])
    ref["randomize_seed1234"] = np.array(draws)
    np.savez_compressed(os.path.join(OUT, "ref_functions.npz"), **ref)
    np.savez_compressed(os.path.join(OUT, "readme_g1.npz"), **g1)

    # ---- G2: sim_data.txt -------------------------------------------------------------------
    with open(os.path.join(REF, "sim_data.txt")) as fh:
        lines = fh.read().strip().splitlines()
    rows = [[float(v) for v in ln.replace(",", " ").split()] for ln in lines[1:] if ln.strip() and not ln.startswith("#")]
    json.dump({"header": lines[0], "rows": rows}, open(os.path.join(OUT, "sim_data_g2.json"), "w"), indent=0)

    # ---- small seeded C2 batch --------------------------------------------------------------
    par = workloads.c2_params(8, 5, seed=7)
    plan = engine.Plan(par, 60.0, 1441, substeps=1, flags=_abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC)
    buf = engine.Buffers(plan)
    oracle.run(plan.cfg, par, buf)
    np.savez_compressed(os.path.join(OUT, "c2_small.npz"), params=par, n_days=60.0, n_outputs=1441,
                        tt_offsets=plan.tt_offsets, ls_offsets=plan.ls_offsets, tt=buf.tt, ttv=buf.ttv, n_tt=buf.n_tt,
                        period=buf.period, t0=buf.t0, ttv_max=buf.ttv_max, mean_a=buf.mean_a, mean_e=buf.mean_e,
                        mean_inc=buf.mean_inc, mean_omega=buf.mean_omega, status=buf.status,
                        ls_oc_power=buf.ls_oc_power, ls_oc_nfreq=buf.ls_oc_nfreq)
    print("golden fixtures written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print("  %-22s %8d bytes" % (fn, os.path.getsize(os.path.join(OUT, fn))))


if __name__ == "__main__":
    main()
