"""Parity statistics of the CUDA path against the CPU oracle over a large seeded sample of the
reference priors (run on the GPU box; TEST INFRASTRUCTURE, uses oracle/).  Writes a JSON summary.

    python tests/parity_sweep.py [n_systems] [out.json]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from nbody_ai_b200 import _abi, engine, workloads  # noqa: E402


def hill_separation(par):
    Ms = par[:, 0, 0]
    d = np.full(len(par), np.inf)
    for j in range(2, par.shape[1]):
        a1, a2 = np.cbrt(par[:, j - 1, 1] ** 2), np.cbrt(par[:, j, 1] ** 2)
        rh = np.cbrt((par[:, j - 1, 0] + par[:, j, 0]) / (3 * Ms)) * 0.5 * (a1 + a2)
        d = np.minimum(d, (a2 - a1) / rh)
    return d


def oracle_is_chaotic(par_b, n_days, n_outputs):
    """The calibration's regularity test (tests/calibrate_substeps.py:is_regular): the oracle's own
    answer must not move by more than 0.01 s in a transit time or 1e-9 in a mean eccentricity when P1 changes
    by 1e-13 (relative).  A system that fails has no defined parity target."""
    P = np.ascontiguousarray(par_b[None])
    plan = engine.Plan(P, n_days, n_outputs, substeps=1, flags=_abi.F_MEANS)
    a = engine.Buffers(plan); oracle.run(plan.cfg, P, a)
    Q = P.copy(); Q[0, 1, 1] *= 1.0 + 1e-13
    b = engine.Buffers(plan); oracle.run(plan.cfg, Q, b)
    if not np.array_equal(a.n_tt, b.n_tt):
        return True
    dtt = max([np.abs(a.tt_of(0, j) - b.tt_of(0, j)).max() * 86400 for j in range(plan.Np) if a.n_transits(0, j)] + [0.0])
    de = np.abs(a.mean_e - b.mean_e) / np.maximum(a.mean_e, 1e-4)
    return bool(dtt > 1e-2 or de.max() > 1e-9)


def sweep(par, n_days, n_outputs, label):
    flags = _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC
    plan = engine.Plan(par, n_days, n_outputs, flags=flags)
    t0 = time.time(); got = engine.run(plan); t_gpu = time.time() - t0
    want = engine.Buffers(plan)
    t0 = time.time(); oracle.run(plan.cfg, plan.params, want); t_cpu = time.time() - t0
    delta = hill_separation(par)
    worst = dict(tt_s=0.0, oc_min=0.0, a_rel=0.0, e_rel=0.0, ls_power=0.0, ttv_max_min=0.0)
    n_cmp = n_count_mismatch = n_chaotic = 0
    fails, top = [], []
    for b in range(plan.B):
        if delta[b] < 4.0 or (want.status[b] & (_abi.S_NONFINITE | _abi.S_UNBOUND | _abi.S_TT_OVERFLOW)):
            continue          # Hill-unstable / broken in the oracle itself: chaotic, no parity defined
        n_cmp += 1
        w = dict(tt_s=0.0, oc_min=0.0, a_rel=0.0, e_rel=0.0, ls_power=0.0, ttv_max_min=0.0)
        for j in range(plan.Np):
            sp = b * plan.Np + j
            if got.n_tt[sp] != want.n_tt[sp]:
                n_count_mismatch += 1
                w['tt_s'] = np.inf
                continue
            n = got.n_transits(b, j)
            if n:
                w['tt_s'] = max(w['tt_s'], np.abs(got.tt_of(b, j) - want.tt_of(b, j)).max() * 86400)
            if n >= 3:
                w['oc_min'] = max(w['oc_min'], np.abs(got.ttv_of(b, j) - want.ttv_of(b, j)).max() * 1440)
                w['ls_power'] = max(w['ls_power'], np.abs(got.ls_oc_of(b, j)[1] - want.ls_oc_of(b, j)[1]).max())
                w['ttv_max_min'] = max(w['ttv_max_min'], abs(got.ttv_max[sp] - want.ttv_max[sp]) * 1440)
            w['a_rel'] = max(w['a_rel'], abs(got.mean_a[sp] / want.mean_a[sp] - 1))
            w['e_rel'] = max(w['e_rel'], abs(got.mean_e[sp] - want.mean_e[sp]) / max(want.mean_e[sp], 1e-4))
        # anything beyond a tenth of a tolerance is checked for regularity first: a chaotic system is excluded
        if (w['tt_s'] > 0.1 or w['oc_min'] > 1e-4 or w['a_rel'] > 1e-9 or w['e_rel'] > 1e-9) and \
                oracle_is_chaotic(par[b], n_days, n_outputs):
            n_chaotic += 1
            n_cmp -= 1
            continue
        if w['tt_s'] > 1.0 or w['oc_min'] > 1e-3 or w['a_rel'] > 1e-8 or w['e_rel'] > 1e-8:
            fails.append(dict(system=int(b), hill_separation=float(delta[b]), P=par[b, 1:, 1].tolist(), **{k: float(v) for k, v in w.items()}))
        for k in worst:
            if np.isfinite(w[k]):
                worst[k] = max(worst[k], w[k])
        top.append((w['e_rel'], int(b)))
    return dict(label=label, n_systems=int(plan.B), n_planets=int(plan.Np), n_days=n_days, n_outputs=n_outputs,
                compared=n_cmp, excluded_hill_unstable_or_broken=int(plan.B - n_cmp - n_chaotic),
                excluded_chaotic_by_oracle_self_sensitivity=n_chaotic, transit_count_mismatches=n_count_mismatch,
                outside_tolerance=len(fails), failures=fails[:10],
                largest_e_rel=[dict(system=b, e_rel=float(e), substeps=int(plan.substeps[b]), P=par[b, 1:, 1].tolist(),
                                    mass_ratio=(par[b, 1:, 0] / par[b, 0, 0]).tolist(), hill_separation=float(delta[b]))
                               for e, b in sorted(top, reverse=True)[:8]], worst={k: float(v) for k, v in worst.items()},
                tolerances=dict(tt_s=1.0, oc_min=1e-3, a_rel=1e-8, e_rel=1e-8),
                gpu_seconds_incl_copies=t_gpu, oracle_seconds=t_cpu, oracle_threads=oracle.max_threads())


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "parity_sweep.json")
    res = [sweep(workloads.random_systems(n, 2, seed=1001), 365.0, 8760, "C2 priors, 2 planets, 365 d @ 1 h"),
           sweep(workloads.random_systems(n // 4, 4, seed=1002), 365.0, 8760, "C5 priors, 4 planets, 365 d @ 1 h")]
    json.dump(res, open(out, "w"), indent=1)
    for r in res:
        print(r["label"], "compared", r["compared"], "outside tolerance", r["outside_tolerance"], "worst", r["worst"])


if __name__ == "__main__":
    main()
