import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench
pkg = g.load_package()
X, lab = bench.synth(10000, 50, 100, 0.1, 50, 44)
import warnings; warnings.simplefilter("ignore")
for algo in ("k-means", "k-medoids"):
    for rep in range(2):
        t = time.perf_counter(); p = pkg.fitprior(X.T, algo, False, Kmax=60, verbose=False, rng=3); dt = time.perf_counter() - t
        print(f"fitprior n=10000 dim=100 {algo} Kmax=60: {dt:.3f} s  K_initial={p.K_initial}", flush=True)
t = time.perf_counter(); r = pkg.kmeans(X.T, 50, rng=1); print("kmeans k=50:", time.perf_counter() - t, r["iterations"], r["converged"], r["totalcost"])
