"""Does overlapping consecutive steps on two CUDA streams fill the tail of the heterogeneous C2 launch?  Device-resident
steps (engine.run_device) alternating between two streams / two buffer sets, against the same steps on one stream."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from nbody_ai_b200 import _abi, engine, workloads  # noqa: E402

fl = _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC
NS = 4
plans = [engine.Plan(workloads.c2_params(10000, 6, seed=s), 365.0, 8760, flags=fl) for s in range(NS)]
bufs = [engine.DeviceBuffers(p) for p in plans]
streams = [torch.cuda.Stream() for _ in range(NS)]
for mode in ("one stream", "two streams", "three streams", "four streams", "two streams, integrate only", "one stream, integrate only"):
    only = "integrate only" in mode
    if only:
        plans2 = [engine.Plan(p.params, 365.0, 8760, flags=_abi.F_MEANS | _abi.F_MEAN_OMEGA) for p in plans]
        bufs2 = [engine.DeviceBuffers(p) for p in plans2]
    P, Bf = (plans2, bufs2) if only else (plans, bufs)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 24
        ns = {"one": 1, "two": 2, "three": 3, "four": 4}[mode.split()[0]]
        for k in range(n):
            with torch.cuda.stream(streams[k % ns]):
                engine.run_device(P[k % ns], Bf[k % ns])
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n * 1e3
    print("%-32s %.2f ms per step" % (mode, dt), flush=True)
