"""Profiling driver: the bench workload (n=10k, 256 chains) with one init launch, one warm-up sweep and
`--steps` profiled sweeps.  Used under ncu (see profiles/README.md); numbers printed here are not bench values."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_stats_lib = os.path.join(ROOT, "redclust.jl_b200", "librcb200_stats.so")
if os.path.exists(_stats_lib):
    os.environ.setdefault("RCB200_LIB", _stats_lib)      # the build with the in-kernel cycle counters
import __graft_entry__ as graft
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10000)
ap.add_argument("--K", type=int, default=50)
ap.add_argument("--dim", type=int, default=100)
ap.add_argument("--chains", type=int, default=256)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--numMH", type=int, default=1)
ap.add_argument("--sigma", type=float, default=0.1)
ap.add_argument("--maxK", type=int, default=0)
ap.add_argument("--warm", type=int, default=1, help="untimed sweeps before the profiled ones")
a = ap.parse_args()
pkg = graft.load_package()
X, lab = bench.synth(a.n, a.K, a.dim, a.sigma, a.K, 44)
data = pkg.MCMCData.from_points(X)
params = pkg.params_from_labels(data.D, lab, maxK=a.maxK)
opts = pkg.MCMCOptionsList(numiters=a.warm + a.steps, burnin=0, thin=1, numMH=a.numMH)
rp = [pkg.init_rp(params, 44, c) for c in range(a.chains)]
smp = pkg.Sampler(data, opts, params, np.tile(lab, (a.chains, 1)), [x[0] for x in rp], [x[1] for x in rp], seed=44)
smp.run(0)
smp.run(a.warm)
st0 = {k: v.copy() for k, v in smp.stats().items()}
t0 = smp.progress()[1]
for _ in range(a.steps):
    smp.run(1)
print('ms per profiled sweep of all chains:', (smp.progress()[1] - t0) / a.steps * 1e3)
print("iters, device seconds:", smp.progress())
st = {k: v - st0[k] for k, v in smp.stats().items()}
its = a.steps
print("per-iteration mean cycles over chains (SM clock):")
for k, v in st.items():
    if k != "-":
        print(f"  {k:20s} {v.mean() / its:14.0f}   min {v.min() / its:14.0f}   max {v.max() / its:14.0f}")
