"""simulations/simulation_grid.py of the reference with the same flags (-e, -n, -o), on the B200 engine:
the whole perturber mass x period grid in one batched call.  Writes ttv_grid.fits with the reference's layout
(:219-295) and an .npz of every grid; the PNG plots (:155-217) are out of scope."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbody_ai_b200.grid import simulation_grid, write_fits  # noqa: E402

if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("-e", "--epochs", help="Number of planet 1 orbital epochs/periods to integrate the simulation for",
                        default=500, type=int)
    parser.add_argument("-n", "--Npts", help="Number of grid points in mass and period ratio", default=150, type=int)
    parser.add_argument("-o", "--n_order", help="Number of harmonics to fit to TTVs", default=4, type=int)
    parser.add_argument("--out", default="simulation_grid.npz")
    parser.add_argument("--fits", default="ttv_grid.fits")
    args = parser.parse_args()
    # astropy.constants values the reference switches to (simulation_grid.py:83-85)
    msun, mearth = 1.988409870698051e30, 5.972167867791379e24
    objects = [
        {'m': 1},
        {'m': 10 * mearth / msun, 'P': 1.42, 'inc': np.deg2rad(90), 'e': 0, 'omega': 0.01},
        {'m': 10 * mearth / msun, 'omega': np.deg2rad(90), 'e': 0, 'inc': np.deg2rad(90)},
    ]
    mass_range = np.linspace(0.5, 325, args.Npts) * mearth / msun
    period_range = np.linspace(1.41, 2.41, args.Npts) * objects[1]['P']
    mass_grid, period_grid = np.meshgrid(mass_range, period_range)
    # the reference fills grid[i][j] from mass_grid[i][j], period_grid[i][j] (:139-141)
    grids = simulation_grid(objects, mass_grid, period_grid, epochs=args.epochs, n_order=args.n_order)
    np.savez_compressed(args.out, mass_grid=mass_grid, period_grid=period_grid, **grids)
    objects[2]['m'], objects[2]['P'] = mass_grid[-1][-1], period_grid[-1][-1]      # what the reference's loop leaves behind (:139-141)
    write_fits(args.fits, grids, mass_grid, period_grid, objects, msun / mearth)
    amp = grids["ttv_grid"] * 24 * 60
    print("grid %dx%d, %d epochs: TTV amplitude %.3f .. %.3f min -> %s" % (args.Npts, args.Npts, args.epochs, np.nanmin(amp), np.nanmax(amp), args.out))
