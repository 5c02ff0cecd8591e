"""Calibrate the sub-step rule of nbt_plan_substeps against the CPU oracle (IAS15).

For random systems drawn from the reference's priors (nbody/simulation.py:54-129) find, per
scheme, the smallest number k of composition steps per output interval for which the host
build of the device integrator (tests/hostsim) meets the parity tolerances with a safety
margin, and print it against P_min / output step.  TEST/CALIBRATION TOOLING: uses oracle/.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from nbody_ai_b200 import _abi, engine  # noqa: E402
from nbody_ai_b200.simulation import randomize  # noqa: E402

hs = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"))
TOL = dict(tt=1.0, oc=1e-3, a=1e-8, e=1e-8)
MARGIN = 5.0


def errors(buf, ob, plan, b=0):
    w = dict(tt=0.0, oc=0.0, a=0.0, e=0.0)
    for j in range(plan.Np):
        x, y = buf.tt_of(b, j), ob.tt_of(b, j)
        if len(x) != len(y):
            return None
        if len(x):
            w['tt'] = max(w['tt'], np.abs(x - y).max() * 86400)
        if len(x) >= 3:
            w['oc'] = max(w['oc'], np.abs(oracle.ttv_fit(x)[0] - oracle.ttv_fit(y)[0]).max() * 1440)
        sp = b * plan.Np + j
        w['a'] = max(w['a'], abs(buf.mean_a[sp] / ob.mean_a[sp] - 1))
        w['e'] = max(w['e'], abs(buf.mean_e[sp] - ob.mean_e[sp]) / max(ob.mean_e[sp], 1e-4))
    return w


def main(n=100, seed=0, scheme='m64', nd=365.0, no=8760):
    rng = np.random.default_rng(seed)
    rows = []
    for i in range(n):
        obj = randomize(rng)
        P = _abi.objects_to_params(obj)[None]
        plan = engine.Plan(P, nd, no, substeps=1, scheme=scheme, flags=_abi.F_MEANS)
        ob = engine.Buffers(plan)
        oracle.run(plan.cfg, P, ob)
        kmin = None
        for k in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32):
            plan.cfg.substeps = k
            buf = engine.Buffers(plan)
            inp, out = buf.as_structs()
            hs.hostsim_run(C.byref(plan.cfg), C.byref(inp), C.byref(out))
            w = errors(buf, ob, plan)
            if w is not None and all(w[q] * MARGIN < TOL[q] for q in TOL):
                kmin = k
                break
        step = nd / (no - 1)
        pmin = P[0, 1:, 1].min()
        mr = P[0, 1:, 0].max() / P[0, 0, 0]
        rows.append((pmin / step, kmin or 99, mr, P[0, 2, 1] / P[0, 1, 1], w))
        print("%3d  Pmin/step %7.1f  kmin %3d  R=%7.1f  massratio %.2e  P2/P1 %.2f  %s" % (
            i, pmin / step, kmin or 99, (kmin or 99) * pmin / step, mr, P[0, 2, 1] / P[0, 1, 1],
            {q: "%.1e" % v for q, v in (w or {}).items()}), flush=True)
    R = np.array([r[0] * r[1] for r in rows])
    print("scheme", scheme, "R needed: max %.1f  p95 %.1f  median %.1f" % (R.max(), np.percentile(R, 95), np.median(R)))


def is_regular(P, nd, no, tol_s=1e-2):
    """Chaos filter: the oracle itself must be insensitive to a 1e-13 relative change of P1."""
    plan = engine.Plan(P, nd, no, substeps=1, flags=_abi.F_MEANS)
    a = engine.Buffers(plan); oracle.run(plan.cfg, P, a)
    Q = P.copy(); Q[0, 1, 1] *= 1.0 + 1e-13
    b = engine.Buffers(plan); oracle.run(plan.cfg, Q, b)
    if not np.array_equal(a.n_tt, b.n_tt):
        return False
    return all(np.abs(a.tt_of(0, j) - b.tt_of(0, j)).max() * 86400 < tol_s for j in range(plan.Np) if a.n_transits(0, j))


def validate(n=200, seed=1, scheme='m64', nd=365.0, no=8760):
    """Planner-chosen k: how many regular systems from the priors meet every tolerance?"""
    rng = np.random.default_rng(seed)
    nfail = nchaos = 0
    ks = []
    for i in range(n):
        obj = randomize(rng)
        P = _abi.objects_to_params(obj)[None]
        plan = engine.Plan(P, nd, no, scheme=scheme, flags=_abi.F_MEANS)
        ob = engine.Buffers(plan); oracle.run(plan.cfg, P, ob)
        buf = engine.Buffers(plan)
        inp, out = buf.as_structs()
        hs.hostsim_run(C.byref(plan.cfg), C.byref(inp), C.byref(out))
        w = errors(buf, ob, plan)
        ok = w is not None and all(w[q] < TOL[q] for q in TOL)
        k = int(plan.substeps[0]); ks.append(k)
        if i % 20 == 0:
            print("  ... %d" % i, flush=True)
        if not ok:
            if not is_regular(P, nd, no):
                nchaos += 1
                print("%3d chaotic (excluded) k=%d P1=%.2f P2/P1=%.2f" % (i, k, P[0, 1, 1], P[0, 2, 1] / P[0, 1, 1]), flush=True)
                continue
            nfail += 1
            print("%3d FAIL k=%d P1=%.2f P2/P1=%.2f mr=%.1e %s" % (i, k, P[0, 1, 1], P[0, 2, 1] / P[0, 1, 1],
                  P[0, 1:, 0].max() / P[0, 0, 0], w), flush=True)
    print("validate %s: %d systems, %d chaotic excluded, %d failures, mean k %.2f, k histogram %s" % (
        scheme, n, nchaos, nfail, np.mean(ks), np.bincount(ks)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "validate":
        validate(n=int(sys.argv[2]) if len(sys.argv) > 2 else 200, scheme=sys.argv[3] if len(sys.argv) > 3 else 'm64')
        sys.exit(0)
    main(n=int(sys.argv[1]) if len(sys.argv) > 1 else 100, scheme=sys.argv[2] if len(sys.argv) > 2 else 'm64')
