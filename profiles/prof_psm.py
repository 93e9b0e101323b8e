"""Profiling driver of the PSM count kernel: n = 10 000 points, 2 000 samples with ~50 clusters (used under ncu)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
os.environ["RCB200_VERBOSE"] = "1"
rng = np.random.default_rng(5)
n, S, K = 10000, 2000, 50
lab = np.sort(rng.integers(1, K + 1, size=n))
L = np.tile(lab, (S, 1)); flip = rng.random(L.shape) < 0.03; L[flip] = rng.integers(1, K + 1, size=int(flip.sum()))
for _ in range(2):
    P = pkg.psm(L)
print("psm diag ok:", bool(np.all(np.diag(P) == 1.0)))
