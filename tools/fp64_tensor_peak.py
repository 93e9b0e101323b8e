"""FP64 tensor-pipe reference for the distance build's roofline: cuBLAS DGEMM (torch.matmul on fp64) at 8192^3, best of
10 -- the denominator `profiles/` quotes for k_gram128 -- then the kernel itself through MCMCData.from_points
(RCB200_VERBOSE prints its CUDA-event time) and generatemixture's oracle co-clustering at the bench size."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench

a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"cuBLAS DGEMM 8192^3: {best:.2f} ms = {2 * 8192 ** 3 / best / 1e9:.1f} TFLOP/s (fp64 tensor peak reference)")
del a, b, c
pkg = g.load_package()
os.environ["RCB200_VERBOSE"] = "1"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
X, lab = bench.synth(n, 50, 100, 0.1, 50, 44)
for mode in (None, "dmma32", "exact"):
    if mode:
        os.environ["RCB200_DISTM"] = mode
    else:
        os.environ.pop("RCB200_DISTM", None)
    pkg.MCMCData.from_points(X)
    t = time.perf_counter(); d = pkg.MCMCData.from_points(X); dt = time.perf_counter() - t
    print(f"MCMCData(points) n={n} dim=100 mode={mode or 'default (k_gram128)'}: {dt * 1e3:.1f} ms host call")
    del d
os.environ.pop("RCB200_DISTM", None)
t = time.perf_counter()
mix = pkg.generatemixture(n, 50, alpha=50, sigma=0.1, dim=100, rng=1)
dt = time.perf_counter() - t
print(f"generatemixture(N={n}, K=50, dim=100) with the oracle co-clustering (5000 draws, 2 n^2 x 250 000 flop Gram): {dt:.2f} s")
