"""End-to-end step time of the C2 bench batch with the three kinds of host storage: torch-pinned BufferPool,
SharedHostPool (POSIX shared memory + cudaHostRegister), pageable numpy.   usage: python tools/e2e_pools_probe.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from nbody_ai_b200 import _abi, engine, workloads  # noqa: E402

par = workloads.c2_params(10000, 6, seed=0)
fl = _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC
pools = {"pinned (torch)": engine.BufferPool(pinned=True), "pageable": engine.BufferPool(pinned=False),
         "shared + cudaHostRegister": engine.SharedHostPool("nbt_probe_%d" % os.getpid(), 480 << 20)}
for name, pool in pools.items():
    for threads in (0, 4):
        ts, tp = [], []
        for k in range(6):
            if hasattr(pool, "reset"):
                pool.reset()
            t0 = time.perf_counter()
            p = engine.Plan(par, 365.0, 8760, flags=fl, pool=pool, threads=threads)
            t1 = time.perf_counter()
            engine.run(p, pool.buffers_for(p))
            t2 = time.perf_counter()
            if k >= 2:
                ts.append(t2 - t1); tp.append(t1 - t0)
        print("%-28s threads=%d  plan %.2f ms  run %.2f ms" % (name, threads, 1e3 * np.mean(tp), 1e3 * np.mean(ts)), flush=True)
    if hasattr(pool, "close"):
        del p
        pool.close()
# raw copies
dev = torch.device("cuda", 0)
n = 340 << 20
src = torch.empty(n, dtype=torch.uint8, device=dev)
a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
sh = engine.SharedHostPool("nbt_probe2_%d" % os.getpid(), n + (1 << 20))
b = torch.from_numpy(sh._take("x", (n,), np.uint8))
for name, dst in (("pinned", a), ("shared+registered", b)):
    for _ in range(2):
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    print("D2H 340 MiB into %-18s %.1f GB/s" % (name, 5 * n / (time.perf_counter() - t0) * 1e-9))
del b
sh.close()
