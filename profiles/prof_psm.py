"""Profiling driver of the PSM count kernel (used under ncu): by default n = 10 000 points, 2 000 samples with ~50 clusters
through psm(); `python profiles/prof_psm.py 30000 4000 100` = device counts only (no host matrix) at a larger size."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
os.environ["RCB200_VERBOSE"] = "1"
rng = np.random.default_rng(5)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 50
lab = np.sort(rng.integers(1, K + 1, size=n))
L = np.tile(lab, (S, 1)); flip = rng.random(L.shape) < 0.03; L[flip] = rng.integers(1, K + 1, size=int(flip.sum()))
if len(sys.argv) > 1:
    import torch
    cnt = torch.empty((n, n), dtype=torch.int32, device="cuda")
    for _ in range(2):
        pkg.psm_counts_dev(L, cnt.data_ptr())
    torch.cuda.synchronize()
    print("diag ok:", bool((torch.diagonal(cnt) == S).all()))
else:
    for _ in range(2):
        P = pkg.psm(L)
    print("psm diag ok:", bool(np.all(np.diag(P) == 1.0)))
