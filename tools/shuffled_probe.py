"""Bench workload with the points in random order (labels not contiguous along the columns: every row tile holds all
~50 labels, ~500 (tile, label) runs instead of ~60): parity of 2 chains x 3 sweeps against the oracle, then throughput."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g, bench
pkg, orc = g.load_package(), g.load_oracle()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 256
X, lab = bench.synth(n, 50, 100, 0.1, 50, 44)
perm = np.random.default_rng(1).permutation(n)
X, lab = X[perm], pkg.sortlabels(lab[perm])
data = pkg.MCMCData.from_points(X)
params = pkg.params_from_labels(data, lab)
P = orc.make_params(**{k: getattr(params, k) for k in params._fields})
init = lab.copy(); idx = np.random.default_rng(2).choice(n, n // 50, replace=False); init[idx] = np.random.default_rng(3).integers(1, 51, size=idx.size)
rp = [pkg.init_rp(params, 5, c) for c in range(2)]
smp = pkg.Sampler(data, pkg.MCMCOptionsList(numiters=3, burnin=0, thin=1), params, np.tile(init, (2, 1)), [a for a, _ in rp], [b for _, b in rp], seed=5)
smp.run(-1)
D = data.D
ok = True
for c in range(2):
    got = smp.samples(c)
    ref = orc.run_chain(D, orc.Options(3, 0, 1, 5, 1), P, init, rp[c][0], rp[c][1], seed=5, chain=c)
    ok &= all(np.array_equal(got[k], ref[k]) for k in ("labels", "K", "r", "p", "loglik", "logposterior", "r_acc", "sm_acc", "sm_split"))
print("shuffled n =", n, "parity with the oracle:", ok, flush=True)
smp.close()
rp = [pkg.init_rp(params, 44, c) for c in range(chains)]
smp = pkg.Sampler(data, pkg.MCMCOptionsList(numiters=8, burnin=0, thin=1), params, np.tile(lab, (chains, 1)), [a for a, _ in rp], [b for _, b in rp], seed=44)
smp.run(3)
_, t0 = smp.progress(); smp.run(5); _, t1 = smp.progress()
print(f"shuffled columns: {chains * 5 / (t1 - t0):.0f} chain-sweeps/s ({(t1 - t0) / 5 * 1e3:.1f} ms per sweep of {chains} chains)", flush=True)
