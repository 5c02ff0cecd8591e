"""simulations/generate_simulations.py of the reference with the same flags (-f, -s, -o, -p), on the
B200 engine: random 2-planet systems -> planet-1 transit times, [X, y] pickle, resume from an existing
file (:27-30)."""
import argparse
import os
import pickle
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbody_ai_b200.dataset import generate_training_set  # noqa: E402

if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("-f", "--file", help="Name of pickle file to save simulations to", default="ttv.pkl")
    parser.add_argument("-s", "--samples", help="Number of simulations", default=100000, type=int)
    parser.add_argument("-o", "--omegas", help="Number of periastron arguments per simulation", default=6, type=int)
    parser.add_argument("-p", "--periods", help="Number of planet 1 orbital periods to integrate the simulation for",
                        default=1000, type=int)
    args = parser.parse_args()
    try:
        X, y = pickle.load(open(args.file, 'rb'))
    except Exception:
        X, y = [], []
    X, y = generate_training_set(args.samples, omegas=args.omegas, periods=args.periods, file=args.file, X=X, y=y)
    print(len(X), "simulations in", args.file)
