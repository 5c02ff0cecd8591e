"""Small end-to-end exercise of every kernel and entry point (all flag combinations, both integrator
kernels, undersized transit capacities).  Written for `compute-sanitizer --tool memcheck`; that tool is
closed on this GPU pool, so it serves as a plain does-everything-run check."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbody_ai_b200 import _abi, engine, grid, nested, simulation, tools, workloads  # noqa: E402

ALL = (_abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_RV | _abi.F_ORBIT | _abi.F_LS_OC | _abi.F_LS_RV)
par = workloads.random_systems(45, 3, seed=5)
for lanes in (0, 45, 7):
    b = engine.run(engine.Plan(par, 40.0, 961, flags=ALL, lanes=lanes))
    assert np.isfinite(b.mean_a).all()
b = engine.run(engine.Plan(par[:5], 20.0, 481, flags=ALL | _abi.F_SERIES, lanes=5))
b = engine.run(engine.Plan(par[:5], 20.0, 481, flags=_abi.F_SERIES | _abi.F_MEANS, lanes=0))
p = engine.Plan(workloads.c1_params(3), 60, 1441, flags=_abi.F_MEANS | _abi.F_LS_OC)
p.tt_offsets[:] = np.arange(10) * 2          # undersized capacities: overflow must stay in bounds
b = engine.run(p)
assert (b.status & _abi.S_TT_OVERFLOW).all()
sim = simulation.generate(workloads.c1_objects(), 30, 721)
ttv = simulation.analyze(sim)
ep, tv, tt = simulation.analyze(sim, ttvfast=True)
tools.maxavg(tv); tools.find_zero(0.0, 1.0, 1.0, -1.0); tools.lomb_scargle(np.arange(50.0), np.sin(np.arange(50.0)))
par4, _ = workloads.c4_params(3)
r = grid.generate_sim_batch(par4, epochs=40, n_order=4)
c3, meta = workloads.c3_params(33)
fwd = nested.get_ttv_batch(c3, ndays=30)
x = np.arange(0, 20, 2.0)
nested.loglike_batch(fwd, 0.5, 0.94, x, 0.5 + 0.94 * x, np.full(len(x), 1e-3))
print("exercise_all_kernels: all kernels ran")
