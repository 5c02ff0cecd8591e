"""GPU probe: device-resident timings of k_integrate for a few batch shapes / flag sets.
usage: python tools/gpu_probe.py [case ...]   (cases: c1, c1means, c1omega, c2, c2k1, c5, peak)"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from nbody_ai_b200 import _abi, engine, workloads  # noqa: E402
from bench import survey_flops as plan_flops  # noqa: E402  (SURVEY 8(d) FLOP model)


def timed(params, nd, no, flags, reps=2, **kw):
    env = os.environ.get("NBT_PROBE_LANES")          # all | none | <int>; default: planner's choice
    if env:
        kw["lanes"] = len(params) if env == "all" else (0 if env == "none" else int(env))
    plan = engine.Plan(params, nd, no, flags=flags, **kw)
    print("  lanes=%d" % plan.cfg.lanes_systems, end="")
    d = engine.DeviceBuffers(plan)
    engine.run_device(plan, d)
    torch.cuda.synchronize()
    ms = np.zeros(4)
    for _ in range(reps):
        engine.run_device(plan, d, timing=True)
        torch.cuda.synchronize()
        ms += engine.stage_timing()
    ms /= reps
    h = d.to_host()
    Np = plan.Np
    fl = plan_flops(plan)
    st = {int(s): int((h.status == s).sum()) for s in np.unique(h.status)}
    print("  B=%d Np=%d steps=%.3g  integrate %.2f ms  fit %.3f ms  ls %.3f ms  rv %.3f ms -> %.3g steps/s  %.2f TF  status %s" % (
        plan.B, Np, plan.n_steps, ms[0], ms[1], ms[2], ms[3], plan.n_steps / (ms[0] * 1e-3), fl / (ms[0] * 1e-3) * 1e-12, st), flush=True)
    return ms


def timed_big(params, nd, no, flags, **kw):
    """One timed pass for a batch too large to copy back in full: status and counts only."""
    t0 = time.time()
    plan = engine.Plan(params, nd, no, flags=flags, **kw)
    d = engine.DeviceBuffers(plan)
    print("  planned in %.1f s: B=%d Np=%d steps=%.3g tt capacity %.3g doubles, k hist %s" % (
        time.time() - t0, plan.B, plan.Np, plan.n_steps, float(plan.tt_offsets[-1]), np.bincount(np.abs(plan.substeps)).tolist()), flush=True)
    engine.run_device(plan, d, timing=True)
    torch.cuda.synchronize()
    ms = engine.stage_timing()
    st = d.status.cpu().numpy()
    ntt = d.n_tt.cpu().numpy()
    fl = plan_flops(plan)
    print("  integrate %.1f ms  fit %.1f ms  ls %.1f ms -> %.3g steps/s  %.2f TF  %.3g sims/s; transits %d; status %s" % (
        ms[0], ms[1], ms[2], plan.n_steps / (ms[0] * 1e-3), fl / (ms[0] * 1e-3) * 1e-12, plan.B / (ms.sum() * 1e-3), int(ntt.sum()),
        {int(s): int((st == s).sum()) for s in np.unique(st)}), flush=True)


def main():
    cases = sys.argv[1:] or ["peak", "c1", "c1means", "c1omega", "c2k1", "c2"]
    lib = _abi.load()
    for c in cases:
        print(c, flush=True)
        if c == "peak":
            tf, ms = C.c_double(), C.c_double()
            lib.nbt_fp64_peak(0, 1 << 16, C.byref(tf), C.byref(ms))
            print("  DFMA peak %.2f TFLOP/s (%.2f ms)" % (tf.value, ms.value))
            lib.nbt_fp64_peak(0, -(1 << 16), C.byref(tf), C.byref(ms))
            print("  DFMA with 3 distinct register operands %.2f TFLOP/s (%.2f ms)" % (tf.value, ms.value))
            l1, l2 = C.c_double(), C.c_double()
            lib.nbt_fp64_latency(0, C.byref(l1), C.byref(l2))
            print("  dependent DFMA latency %.1f cycles; MUFU.RSQ64H+DADD %.1f cycles" % (l1.value, l2.value))
        elif c == "c1":
            timed(workloads.c1_params(65536), 30.0, 721, 0, substeps=1)
        elif c == "c1means":
            timed(workloads.c1_params(65536), 30.0, 721, _abi.F_MEANS, substeps=1)
        elif c == "c1omega":
            timed(workloads.c1_params(65536), 30.0, 721, _abi.F_MEANS | _abi.F_MEAN_OMEGA, substeps=1)
        elif c == "c1small":
            timed(workloads.c1_params(16384), 10.0, 241, _abi.F_MEANS, substeps=1, reps=1)
        elif c == "c2k1":
            timed(workloads.c2_params(10000, 6, 0), 30.0, 721, _abi.F_MEANS, substeps=1)
        elif c == "c2":
            timed(workloads.c2_params(10000, 6, 0), 30.0, 721, _abi.F_MEANS)
        elif c == "c2full":
            timed(workloads.c2_params(10000, 6, 0), 365.0, 8760, _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC, reps=1)
        elif c.startswith("c2fullcap"):
            cap = int(c[len("c2fullcap"):])
            par = workloads.c2_params(10000, 6, 0)
            pl = engine.Plan(par, 365.0, 8760, flags=_abi.F_MEANS)
            timed(par, 365.0, 8760, _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC, reps=1, substeps=np.minimum(pl.substeps, cap))
        elif c.startswith("c2heavy"):
            par = workloads.c2_params(10000, 6, 0)
            pl = engine.Plan(par, 365.0, 8760, flags=_abi.F_MEANS)
            sel = par[pl.substeps == 3]
            fl = {"c2heavy0": 0, "c2heavy1": _abi.F_MEANS, "c2heavy2": _abi.F_MEANS | _abi.F_MEAN_OMEGA,
                  "c2heavy3": _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC, "c2heavyp1": _abi.F_PLANET1_ONLY}[c]
            timed(sel, 365.0, 8760, fl, reps=1, substeps=3)
        elif c == "c5full":      # BASELINE config 5: 1M random 4-planet systems, 3-year integrations
            timed_big(workloads.c5_params(1000000), 1095.0, 26280, _abi.F_MEANS)
        elif c == "c4full":      # BASELINE config 4: 256 x 256 grid, 500 epochs of a 1.42 d planet at 30 min
            par, meta = workloads.c4_params(256)
            timed_big(par, meta["ndays"], meta["noutputs"], _abi.F_PLANET1_ONLY)
        elif c == "c1rv":      # full analyze() extras: RV series, its maxavg and its periodogram (8759 x 1816)
            timed(workloads.c1_params(4096), 365.0, 8760, _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC | _abi.F_RV | _abi.F_LS_RV, reps=1, lanes=0)
        elif c in ("c2clean", "c2cleancap1", "c2dirty"):
            # the C2 batch without / only its Hill-unstable systems (mutual separation < 4 R_Hill)
            par = workloads.c2_params(10000, 6, 0)
            a1, a2 = np.cbrt(par[:, 1, 1] ** 2), np.cbrt(par[:, 2, 1] ** 2)
            rh = np.cbrt((par[:, 1, 0] + par[:, 2, 0]) / (3 * par[:, 0, 0])) * 0.5 * (a1 + a2)
            clean = (a2 - a1) / rh >= 4.0
            sel = par[~clean] if c == "c2dirty" else par[clean]
            fl = _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC
            if c == "c2cleancap1":
                timed(sel, 365.0, 8760, fl, reps=1, substeps=1)
            else:
                timed(sel, 365.0, 8760, fl, reps=1)
        elif c in ("c2full_nomeans", "c2full_means"):   # the bench batch without / with part of the per-sample work
            fl = {"c2full_nomeans": _abi.F_LS_OC, "c2full_means": _abi.F_MEANS | _abi.F_LS_OC}[c]
            timed(workloads.c2_params(10000, 6, 0), 365.0, 8760, fl, reps=1)
        elif c == "c1big":
            timed(workloads.c1_params(227328), 30.0, 721, 0, substeps=1)
        elif c == "c1bigomega":
            timed(workloads.c1_params(227328), 30.0, 721, _abi.F_MEANS | _abi.F_MEAN_OMEGA, substeps=1)
        elif c == "c2big":
            par = workloads.c2_params(40000, 6, 0)
            par = par[(par[:, 1, 1] > 2.6) & (par[:, 2, 1] / par[:, 1, 1] > 2.0)][:227328]
            timed(par, 30.0, 721, _abi.F_MEANS | _abi.F_MEAN_OMEGA, substeps=1)
        elif c.startswith("c2d"):        # the bench workload at another size: <draws> x 7 systems
            timed(workloads.c2_params(int(c[3:]), 6, 0), 365.0, 8760, _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC, reps=1)
        elif c.startswith("c2seed"):     # the bench batch of another seed (rank r of a multi-GPU run before balancing)
            timed(workloads.c2_params(10000, 6, int(c[len("c2seed"):])), 365.0, 8760, _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC, reps=1)
        elif c == "c5":
            timed(workloads.c5_params(65536), 30.0, 721, _abi.F_MEANS)
        elif c == "c5lanes":     # the same on the one-planet-per-lane kernel
            timed(workloads.c5_params(65536), 30.0, 721, _abi.F_MEANS, lanes=65536)
        elif c == "c1biglanes":
            timed(workloads.c1_params(227328), 30.0, 721, 0, substeps=1, lanes=227328)
        elif c == "c3":
            par, meta = workloads.c3_params(5000)
            timed(par, meta["ndays"], meta["noutputs"], _abi.F_PLANET1_ONLY, reps=3)
        elif c == "c1one":
            timed(workloads.c1_params(1), 365.0, 8760, _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC, reps=3)


if __name__ == "__main__":
    main()
