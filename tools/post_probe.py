"""Where the time goes in fitprior and in the configs[4]-shaped PSM (host-side breakdowns)."""
import cProfile, pstats, os, sys, time, io
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench
pkg = g.load_package()
X, lab = bench.synth(10000, 50, 100, 0.1, 50, 44)
data = pkg.MCMCData.from_points(X)
pkg.fitprior(data, "k-medoids", True, Kmin=1, Kmax=10, verbose=False, rng=1)
pr = cProfile.Profile(); pr.enable()
t = time.perf_counter(); p = pkg.fitprior(data, "k-medoids", True, Kmin=1, Kmax=60, verbose=False, rng=1); dt = time.perf_counter() - t
pr.disable()
print(f"fitprior(device, k-medoids, K=1..60): {dt:.2f} s -> K_initial={p.K_initial}")
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14); print(s.getvalue()[-2600:])
del data
os.environ["RCB200_VERBOSE"] = "1"
import torch
gg = np.random.default_rng(1)
n2, S2 = 50000, 10000
base = np.sort(gg.integers(1, 91, size=n2))
t = time.perf_counter()
L2 = np.tile(base, (S2, 1)); flip = gg.random((S2, n2)) < 0.15; L2[flip] = gg.integers(1, 101, size=int(flip.sum())); L2 = np.ascontiguousarray(L2, dtype=np.int64)
print(f"label generation {time.perf_counter() - t:.1f} s", flush=True)
cnt = torch.empty((n2, n2), dtype=torch.int32, device="cuda")
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); pkg.psm_counts_dev(L2, cnt.data_ptr()); torch.cuda.synchronize()
    print(f"psm_counts_dev n={n2} S={S2}: {time.perf_counter() - t:.2f} s", flush=True)
# MPEL at the configs[4] scale: S = 10 000 candidate samples of n = 50 000 (5e7 pairs of contingency tables)
if len(sys.argv) > 1 and sys.argv[1] == "mpel":
    for loss in ("binder", "VI"):
        t = time.perf_counter(); sums, best = pkg.mpel_loss_sums(L2, loss); dt = time.perf_counter() - t
        print(f"getpointestimate(MPEL, {loss}) search n={n2} S={S2}: {dt:.2f} s ({S2 * (S2 - 1) / 2 / dt:.3e} pairs/s), best sample {best}, loss sum {sums[best]:.6g}", flush=True)
