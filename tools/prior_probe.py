"""Timing of the device fitprior pieces at the bench size (n = 10 000): k-medoids per K, pair statistics, and the
host numpy k-medoids for comparison at one K."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g, bench
pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
X, lab = bench.synth(n, 50, 100, 0.1, 50, 44)
data = pkg.MCMCData.from_points(X)
for k in (2, 10, 50, 100):
    t = time.perf_counter(); r = pkg.kmedoids(data, k, rng=1); dt = time.perf_counter() - t
    print(f"device k-medoids n={n} k={k}: {dt * 1e3:.1f} ms, {r['iterations']} iterations, converged={r['converged']}, cost={r['totalcost']:.4f}", flush=True)
t = time.perf_counter(); st = pkg.pair_stats(data, lab); dt = time.perf_counter() - t
print(f"device pair_stats: {dt * 1e3:.1f} ms  nA={st['nA']} nB={st['nB']}", flush=True)
t = time.perf_counter(); p = pkg.params_from_labels(data, lab); print(f"params_from_labels(device): {(time.perf_counter() - t) * 1e3:.1f} ms")
D = data.D
t = time.perf_counter(); p2 = pkg.params_from_labels(D, lab); print(f"params_from_labels(host numpy): {(time.perf_counter() - t) * 1e3:.1f} ms")
t = time.perf_counter(); r = pkg.kmedoids(D, 50, rng=1); dt = time.perf_counter() - t
print(f"host numpy k-medoids k=50: {dt * 1e3:.1f} ms converged={r['converged']}", flush=True)
t = time.perf_counter(); pr = pkg.fitprior(data, "k-medoids", True, Kmin=1, Kmax=60, verbose=False, rng=1); dt = time.perf_counter() - t
print(f"fitprior(device, k-medoids, K=1..60): {dt:.2f} s -> K_initial={pr.K_initial}", flush=True)
