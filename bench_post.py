#!/usr/bin/env python
"""Timings of the pieces around the sampler (BASELINE metric: "PSM + pointestimate secs"): distance-matrix build
(exact fp64 kernel vs FP64 tensor-core DMMA), posterior similarity matrix, MPEL point estimate.  Wall-clock through
the public host API (host buffers in, host buffers out).  Prints one JSON line per piece."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10000)
ap.add_argument("--dim", type=int, default=100)
ap.add_argument("--K", type=int, default=50)
ap.add_argument("--S", type=int, default=2000)
ap.add_argument("--S-mpel", type=int, default=1000)
a = ap.parse_args()
pkg = graft.load_package()
X, lab = bench.synth(a.n, a.K, a.dim, 0.1, a.K, 44)

def timed(f, reps=3):
    f(); ts = []
    for _ in range(reps):
        t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
    return min(ts)

for mode in ("exact", "dmma"):
    os.environ["RCB200_DISTM"] = mode
    t = timed(lambda: pkg.MCMCData.from_points(X))
    print(json.dumps({"piece": "MCMCData(points): distM + logD + fixed-point images", "mode": mode, "n": a.n, "dim": a.dim,
                      "seconds": t, "gflops_distm_only": None}))
os.environ.pop("RCB200_DISTM")
rng = np.random.default_rng(0)
def perturbed(S):
    L = np.tile(lab, (S, 1))
    flip = rng.random(L.shape) < 0.03
    L[flip] = rng.integers(1, a.K + 1, size=int(flip.sum()))
    return L
L = perturbed(a.S)
t = timed(lambda: pkg.psm(L), reps=2)
print(json.dumps({"piece": "PSM (host labels -> host n x n fp64)", "n": a.n, "S": a.S, "seconds": t,
                  "label_compares_per_s": a.n * a.n * a.S / t}))
Lm = L[:a.S_mpel]
for loss in ("binder", "VI"):
    t = timed(lambda: pkg.mpel_loss_sums(Lm, loss), reps=2)
    print(json.dumps({"piece": "MPEL getpointestimate", "loss": loss, "n": a.n, "S": a.S_mpel, "seconds": t,
                      "pairs_per_s": a.S_mpel * (a.S_mpel - 1) / 2 / t}))
