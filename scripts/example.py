"""The reference's example.py (:1-19) on the B200 engine: one system, 1000 inner orbits at 1-hour
cadence, generate() + analyze(); prints the summary numbers report() would tabulate."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbody_ai_b200.simulation import analyze, generate  # noqa: E402
from nbody_ai_b200.tools import mearth, mjup, msun  # noqa: E402

if __name__ == "__main__":
    objects = [
        {'m': 0.95},
        {'m': 1.169 * mjup / msun, 'P': 2.797436, 'inc': 3.14159 / 2, 'e': 0, 'omega': 0},
        {'m': 0.1 * mjup / msun, 'P': 2.797436 * 1.9, 'inc': 3.14159 / 2, 'e': 0.0, 'omega': 0},
    ]  # HAT-P-37
    t1 = time.time()
    n_orbits = 1000
    sim_data = generate(objects, objects[1]['P'] * n_orbits, int(n_orbits * objects[1]['P'] * 24))
    print("Simulation time: {}s".format(time.time() - t1))
    ttv_data = analyze(sim_data)
    print("RV max-avg [m/s]: %.2f" % ttv_data['RV']['max'])
    for j, p in enumerate(ttv_data['planets']):
        print("planet %d: mass %.2f Mearth  a %.4f au  P %.5f d  e %.4f  inc %.2f deg  %d transits  TTV max-avg %.2f min" % (
            j + 1, p['mass'] * msun / mearth, p['a'], p['P'], p['e'], p['inc'] * 180 / 3.141592653589793, len(p['tt']), p['max'] * 24 * 60))
