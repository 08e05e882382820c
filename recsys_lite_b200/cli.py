"""``python -m recsys_lite_b200 benchmark ratings.csv --rank K --top-n N --iterations I [--algorithm svd|als|topn]``

The B200 version of the reference's ``recsys benchmark`` command (cli.py:179-211 -> bench :153-176): same options,
same ``Time | RMSE | MAE`` line per iteration, plus ``--algorithm`` so ALS and top-N are reachable (north star).
The ratings file is a dense users x items CSV (what the reference's ``load_ratings`` reads for .csv) or a scipy
``.npz`` sparse matrix."""
from __future__ import annotations

import argparse
import sys

import numpy as np
from scipy import sparse


def load_matrix(path: str):
    if path.endswith(".npz"):
        return sparse.load_npz(path).tocsr()
    try:
        return np.loadtxt(path, delimiter=",", dtype=np.float32, ndmin=2)
    except ValueError:                       # header line
        return np.loadtxt(path, delimiter=",", dtype=np.float32, ndmin=2, skiprows=1)


def bench(path: str, k: int = 50, n: int = 10, iterations: int = 1, algorithm: str = "svd") -> int:
    from .benchmark import quick_benchmark
    failures = 0
    for i in range(iterations):
        print(f"\nIteration {i + 1}/{iterations}")
        ratings = load_matrix(path)          # the reference reloads the file every iteration (cli.py:163)
        result = quick_benchmark(ratings, k=k, n_runs=1, algorithm=algorithm, n=n)
        if result.get("time") is None:
            print(f"Benchmark failed for k={k}: {result.get('error', 'unknown error')}")
            failures += 1
            continue
        print(f"Time: {result['time']:.4f}s | RMSE: {result['rmse']:.4f} | MAE: {result['mae']:.4f}")
    return failures


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="recsys_lite_b200")
    sub = ap.add_subparsers(dest="command", required=True)
    b = sub.add_parser("benchmark", help="Run performance benchmark on the recommender system")
    b.add_argument("csv")
    b.add_argument("--rank", "-k", type=int, default=50)
    b.add_argument("--top-n", "-n", type=int, default=10)
    b.add_argument("--iterations", "-i", type=int, default=1)
    b.add_argument("--algorithm", "-a", default="svd", choices=["svd", "als", "topn"])
    args = ap.parse_args(argv)
    try:
        print(f"Running benchmark with {args.iterations} iteration(s)...")
        failures = bench(args.csv, k=args.rank, n=args.top_n, iterations=args.iterations, algorithm=args.algorithm)
        if args.iterations > 1:
            print(f"\nCompleted {args.iterations} benchmark iterations")
        return 1 if failures else 0
    except Exception as e:                   # cli.py:209-211
        print(f"Benchmark error: {e}")
        return 1


if __name__ == "__main__":
    sys.exit(main())
