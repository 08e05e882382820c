"""``python -m recsys_lite_b200 benchmark ratings.csv --rank K --top-n N --iterations I [--algorithm svd|als|topn]``
``python -m recsys_lite_b200 predict ratings.{csv,npz} --rank K --top-n N [--algorithm svd|als] [--output recs.csv] [--metrics]``

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


def predict(path: str, k: int = 50, n: int = 10, output=None, show_metrics: bool = False, max_users: int = 5,
            algorithm: str = "svd", iterations: int = 10) -> np.ndarray:
    """The reference's ``recsys predict`` (cli.py:34-150): load, reconstruct, top-N, optional RMSE / MAE and CSV output.
    ``--algorithm svd`` is the reference's pipeline svd_reconstruct -> top_n; ``--algorithm als`` routes to the fused path
    RecommenderSystem("als").fit(...).recommend(...) -- a scipy ``.npz`` pattern is never densified there."""
    from . import RecommenderSystem, compute_mae, compute_rmse, svd_reconstruct, top_n
    ratings = load_matrix(path)
    print(f"Loaded {ratings.shape[0]} users, {ratings.shape[1]} items")
    recon = model = None
    if algorithm == "als":
        model = RecommenderSystem("als").fit(ratings, k=k, iterations=iterations)
        recs = model.recommend(ratings, n=n)
    else:
        dense = ratings.toarray() if sparse.issparse(ratings) else ratings
        recon = svd_reconstruct(dense, k=k, use_sparse=False)
        recs = top_n(recon, dense, n=n)
    print(f"\nGenerated recommendations for {len(recs)} users")
    for i, row in enumerate(recs[:max_users]):
        print(f"User {i}: " + ", ".join(str(int(x)) for x in row if x >= 0))
    if show_metrics and model is not None:       # from the resident factors over the observed cells: nothing is densified
        print(f"RMSE: {compute_rmse(model, ratings):.4f}\nMAE: {compute_mae(model, ratings):.4f}")
    elif show_metrics:
        dense = ratings.toarray() if sparse.issparse(ratings) else ratings
        print(f"RMSE: {compute_rmse(recon, dense):.4f}\nMAE: {compute_mae(recon, dense):.4f}")
    if output:
        rec_matrix = np.zeros((len(recs), n), dtype=np.float32)          # the reference's output format (cli.py:141-146)
        rec_matrix[:, : recs.shape[1]] = np.where(recs >= 0, recs, 0)
        np.savetxt(output, rec_matrix, delimiter=",", fmt="%g")
        print(f"Saved recommendations to {output}")
    return recs


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="recsys_lite_b200")
    sub = ap.add_subparsers(dest="command", required=True)
    pr = sub.add_parser("predict", help="Generate top-N recommendations for users")
    pr.add_argument("csv")
    pr.add_argument("--rank", "-k", type=int, default=50)
    pr.add_argument("--top-n", "-n", type=int, default=10)
    pr.add_argument("--output", "-o", default=None)
    pr.add_argument("--metrics", "-m", action="store_true")
    pr.add_argument("--max-users", type=int, default=5)
    pr.add_argument("--algorithm", "-a", default="svd", choices=["svd", "als"])
    pr.add_argument("--iterations", "-i", type=int, default=10)
    b = sub.add_parser("benchmark", help="Run performance benchmark on the recommender system")
    b.add_argument("csv")
    b.add_argument("--rank", "-k", type=int, default=50)
    b.add_argument("--top-n", "-n", type=int, default=10)
    b.add_argument("--iterations", "-i", type=int, default=1)
    b.add_argument("--algorithm", "-a", default="svd", choices=["svd", "als", "topn"])
    args = ap.parse_args(argv)
    if args.command == "predict":
        try:
            predict(args.csv, k=args.rank, n=args.top_n, output=args.output, show_metrics=args.metrics, max_users=args.max_users,
                    algorithm=args.algorithm, iterations=args.iterations)
            return 0
        except Exception as e:               # cli.py:148-150
            print(f"Error: {e}")
            return 1
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
