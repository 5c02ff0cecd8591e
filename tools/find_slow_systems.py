"""Which warps / systems of a bench batch are the critical path of k_integrate?  Times every warp of the
planned launch order that holds a Hill-unstable or multi-step system on its own (a lone warp: pure latency),
then every system of the slowest warps.  Run on the GPU box:  python tools/find_slow_systems.py [seed]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from nbody_ai_b200 import _abi, engine, workloads  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 2
par = workloads.c2_params(10000, 6, seed)
flags = _abi.F_MEANS | _abi.F_MEAN_OMEGA
plan = engine.Plan(par, 365.0, 8760, flags=flags)
order = plan.order if plan.order is not None else np.arange(plan.B)
a1, a2 = np.cbrt(par[:, 1, 1] ** 2), np.cbrt(par[:, 2, 1] ** 2)
rh = np.cbrt((par[:, 1, 0] + par[:, 2, 0]) / (3 * par[:, 0, 0])) * 0.5 * (a1 + a2)
delta = (a2 - a1) / rh


def solo(idx):
    p = engine.Plan(par[idx], 365.0, 8760, substeps=plan.substeps[idx], flags=flags, sort=False, lanes=0)
    d = engine.DeviceBuffers(p)
    engine.run_device(p, d, timing=True)
    torch.cuda.synchronize()
    return engine.stage_timing()[0], d.status.cpu().numpy()


rows = []
for w in range(0, plan.B // 32):
    idx = order[32 * w:32 * w + 32]
    if not ((delta[idx] < 4.0).any() or (np.abs(plan.substeps[idx]) > 1).any()):
        continue
    ms, st = solo(idx)
    rows.append((ms, w))
rows.sort(reverse=True)
print("seed %d: %d warps timed alone; slowest:" % (seed, len(rows)))
for ms, w in rows[:6]:
    idx = order[32 * w:32 * w + 32]
    print("  warp %4d  %.1f ms   k %s  unstable %d" % (w, ms, sorted(set(np.abs(plan.substeps[idx]).tolist())), int((delta[idx] < 4).sum())))
for ms, w in rows[:2]:
    idx = order[32 * w:32 * w + 32]
    per = []
    for b in idx:
        m1, st = solo(np.array([b]))
        per.append((m1, int(b), int(st[0])))
    per.sort(reverse=True)
    print("  warp %d systems (ms alone, index, status, k, Delta, P1, P2, m1/M, m2/M, e1, e2):" % w)
    for m1, b, st in per[:5]:
        print("    %.1f  %6d  st %2d  k %d  D %.2f  P %.3f %.3f  m %.1e %.1e  e %.3f %.3f" % (
            m1, b, st, plan.substeps[b], delta[b], par[b, 1, 1], par[b, 2, 1], par[b, 1, 0] / par[b, 0, 0], par[b, 2, 0] / par[b, 0, 0],
            par[b, 1, 2], par[b, 2, 2]))
