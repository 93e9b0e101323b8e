import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench
pkg = g.load_package()
from redclust_jl_b200 import prior
X, lab = bench.synth(10000, 50, 100, 0.1, 50, 44)
import warnings; warnings.simplefilter("ignore")
def T(f):
    t = time.perf_counter(); r = f(); return r, time.perf_counter() - t
for rep in range(4):
    dev, t0 = T(lambda: pkg.MCMCData.from_points(X))
    rng = np.random.default_rng(3)
    ts = []
    for k in range(1, 61):
        r, t = T(lambda: prior.kmedoids(dev, k, rng=rng)); ts.append((t, r["iterations"]))
    t1 = sum(t for t, _ in ts)
    its = sum(i for _, i in ts)
    r, t2 = T(lambda: pkg.pair_stats(dev, r["assignments"]))
    sizes = np.bincount(lab)[1:]
    r, t3 = T(lambda: prior.sample_rp(sizes, rng=rng))
    tk = []
    for k in range(1, 61):
        r, t = T(lambda: prior.kmeans(X.T, k, rng=rng)); tk.append((t, r["iterations"]))
    print(f"rep {rep}: from_points {t0:.3f}  kmedoids x60 {t1:.3f} ({its} its, max {max(ts)[0]:.3f})  pair_stats {t2:.3f}  sample_rp {t3:.3f}  kmeans x60 {sum(t for t,_ in tk):.3f} ({sum(i for _,i in tk)} its)", flush=True)
    del dev
