"""Randomised parity stress of the sampler against the CPU oracle: many small problems of varying separation, start state and
options (the exact shortcuts -- row summaries, merge bound -- fire on the separated ones and must not change a bit).
usage: python tools/stress_parity.py [cases] [seed] [max n]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

pkg = g.load_package(); orc = g.load_oracle()
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
nmax = int(sys.argv[3]) if len(sys.argv) > 3 else 420
bad = 0; fast_total = 0; quick_total = 0; moved_total = 0
t0 = time.time()
for case in range(cases):
    n = int(rng.integers(60, nmax)); K = int(rng.integers(2, 13)); dim = int(rng.integers(2, 14)); sig = float(rng.choice([0.05, 0.1, 0.15, 0.25, 0.4, 0.6]))
    w = rng.dirichlet(np.full(K, float(K)))
    lab = np.sort(rng.choice(K, size=n, p=w)) + 1
    lab = (np.unique(lab, return_inverse=True)[1] + 1).astype(np.int64)
    X = rng.normal(0, sig, size=(n, max(dim, K))); X[np.arange(n), lab - 1] += 1.0
    perm = rng.permutation(n) if rng.random() < 0.5 else np.arange(n)
    X, lab = X[perm], lab[perm]
    data = pkg.MCMCData.from_points(X); D = data.D
    params = pkg.params_from_labels(D, lab, maxK=int(rng.choice([0, 0, K + 3, 40])) if sig < 0.2 else int(rng.choice([K + 3, 40, 60])))   # (loose mixtures open hundreds of clusters)
    params.repulsion = bool(rng.random() < 0.8)
    mode = rng.integers(0, 3)
    init = lab.copy() if mode == 0 else (np.ones(n, np.int64) if mode == 1 else rng.integers(1, K + 2, size=n))
    init = (np.unique(init, return_inverse=True)[1] + 1).astype(np.int64)
    iters = int(rng.integers(15, 50)); numGibbs = int(rng.integers(0, 6)); numMH = int(rng.choice([1, 1, 1, 2, 3])); nch = int(rng.integers(1, 5))
    burn = int(rng.integers(0, 5)); thin = int(rng.integers(1, 3))
    seed = int(rng.integers(0, 2**31))
    opts = pkg.MCMCOptionsList(numiters=iters, burnin=burn, thin=thin, numGibbs=numGibbs, numMH=numMH)
    rp = [pkg.init_rp(params, seed, c) for c in range(nch)]
    smp = pkg.Sampler(data, opts, params, np.tile(init, (nch, 1)), [a for a, _ in rp], [b for _, b in rp], seed=seed, slot_cap=255)
    smp.run(-1)
    ok = smp.check_sums() == (0, 0)
    P = orc.make_params(**{k: getattr(params, k) for k in params._fields})
    st_ = smp.stats(); fast_total += int(st_["bulk_rows"].sum()); quick_total += int(st_["dec_wait"].sum())
    for c in range(nch):
        got = smp.samples(c)
        ref = orc.run_chain(D, orc.Options(iters, burn, thin, numGibbs, numMH), P, init, rp[c][0], rp[c][1], seed=seed, chain=c)
        for k in ("r_acc", "sm_split", "sm_acc", "K", "labels", "r", "p", "loglik", "logposterior"):
            if not np.array_equal(got[k], ref[k]): ok = False; print(f"case {case} chain {c}: {k} differs", flush=True)
        st = smp.state(c)
        if not np.array_equal(st.clusts, ref["final_labels"]): ok = False; print(f"case {case} chain {c}: final labels differ", flush=True)
        moved_total += int((np.diff(ref["labels"].astype(np.int64), axis=0) != 0).sum())
    bad += not ok
    smp.close()
print(f"{cases} cases, {bad} with differences; rows decided by summaries {fast_total}, merge proposals rejected by the bound {quick_total}, "
      f"label changes between recorded samples {moved_total}; {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
