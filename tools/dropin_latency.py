"""Wall-clock latency of the drop-in calls for one README system (generate + analyze, get_ttv)."""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nbody_ai_b200.simulation import generate, analyze
from nbody_ai_b200 import workloads, nested
obj = workloads.c1_objects()
generate(obj, 30, 721)   # warm-up (CUDA context, module load)
for _ in range(3):
    t0 = time.perf_counter(); sd = generate(obj, 365, 8760); t1 = time.perf_counter(); td = analyze(sd); t2 = time.perf_counter()
    print("generate %.1f ms  analyze %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
t0 = time.perf_counter(); r = nested.get_ttv(obj, ndays=79); t1 = time.perf_counter()
print("get_ttv(ndays=79) %.1f ms, %d transits" % ((t1 - t0) * 1e3, len(r[2])))
