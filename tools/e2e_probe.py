"""Host-pointer nbt_run (the e2e leg of bench.py) per scheme: wall time per call and the in-stream
stage times.  Run on the GPU box:  python tools/e2e_probe.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402,F401  (CUDA context, pinned memory)

from nbody_ai_b200 import _abi, engine, workloads  # noqa: E402

par = workloads.c2_params(10000, 6, seed=0)
flags = _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC
for scheme in sys.argv[1:] or ["sbabc3", "auto"]:
    plan = engine.Plan(par, 365.0, 8760, flags=flags | _abi.F_TIMING, scheme=scheme)
    plan.pin()
    buf = engine.Buffers(plan, pinned=True)
    for _ in range(2):
        engine.run(plan, buf)
    t = []
    for _ in range(4):
        t0 = time.perf_counter()
        engine.run(plan, buf)
        t.append((time.perf_counter() - t0) * 1e3)
    print(scheme, "alt_systems", plan.cfg.alt_systems, "wall ms", np.round(t, 2), "stages", np.round(engine.stage_timing(), 2), flush=True)

if os.environ.get("E2E_PROBE_DEVICE_FIRST"):
    # the bench's sequence: device-resident runs first, then the host-pointer calls
    plan = engine.Plan(par, 365.0, 8760, flags=flags)
    plan.pin()
    buf = engine.Buffers(plan, pinned=True)
    d = engine.DeviceBuffers(plan)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for k in range(8):
        flush.fill_(k)
        engine.run_device(plan, d, timing=True)
        torch.cuda.synchronize()
    print("device stages", np.round(engine.stage_timing(), 2))
    for rep in range(6):
        t0 = time.perf_counter()
        engine.run(plan, buf)
        print("  host call %.2f ms" % ((time.perf_counter() - t0) * 1e3), "stages", np.round(engine.stage_timing(), 2), flush=True)
