"""BASELINE configs[1] as a replay check: n = 1000, K = 20, dim = 50, 64 chains x 10 000 iterations on one B200; chains 0 and
37 are re-run on the CPU oracle (same structured random stream) and every recorded sample is compared bit for bit."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g, bench
pkg, orc = g.load_package(), g.load_oracle()
n, K, dim, chains, iters = 1000, 20, 50, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 10000
X, lab = bench.synth(n, K, dim, 0.2, K, 44)                      # sigma 0.2: boundary points keep moving
data = pkg.MCMCData.from_points(X)
params = pkg.params_from_labels(data, lab)
opts = pkg.MCMCOptionsList(numiters=iters, burnin=iters // 5, thin=10)
rp = [pkg.init_rp(params, 44, c) for c in range(chains)]
smp = pkg.Sampler(data, opts, params, np.tile(lab, (chains, 1)), [a for a, _ in rp], [b for _, b in rp], seed=44)
t = time.perf_counter(); smp.run(-1); wall = time.perf_counter() - t
_, dev = smp.progress()
print(f"GPU: {chains} chains x {iters} iterations in {dev:.2f} s device time ({chains * iters / dev:.0f} chain-sweeps/s), wall {wall:.2f} s", flush=True)
D = data.D
P = orc.make_params(**{k: getattr(params, k) for k in params._fields})
ok = True
for c in (0, 37):
    got = smp.samples(c)
    t = time.perf_counter()
    ref = orc.run_chain(D, orc.Options(iters, iters // 5, 10, 5, 1), P, lab, rp[c][0], rp[c][1], seed=44, chain=c)
    same = all(np.array_equal(got[k], ref[k]) for k in ("labels", "K", "r", "p", "loglik", "logposterior", "r_acc", "sm_acc", "sm_split"))
    moved = int((np.diff(ref["labels"], axis=0) != 0).sum())
    print(f"chain {c}: oracle {time.perf_counter() - t:.1f} s; {ref['labels'].shape[0]} samples, K in [{ref['K'].min()}, {ref['K'].max()}], "
          f"{moved} label changes between consecutive samples, split-merge acceptance {ref['sm_acc'].mean():.4f}; bit-identical: {same}", flush=True)
    ok &= same
print("REPLAY OK" if ok else "REPLAY MISMATCH")
